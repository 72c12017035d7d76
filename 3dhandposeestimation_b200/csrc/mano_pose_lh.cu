// mano_pose_lh.cu — the per-hand "pose stage" of the MANO layer with LANE = HAND (large batches).
//
//   forward :  PCA coefficients -> axis-angle (MANOLayer.py:126-128), 17 Rodrigues (:82-112), pose
//              feature (:114-120), folded joint regression (:139-141), kinematic chain (:159-165),
//              rest-pose removal with the global rotation folded in (:169-175, :188, :204-205)
//   backward:  SURVEY Appendix A.2 steps 3-7 (reverse chain, Rodrigues backward, PCA^T).
//
// The one-warp-per-hand kernels of mano_pose.cu (the mapping north_star names) spend 1.5 k / 2.9 k
// warp-instructions per hand [round-1 ncu: issue-bound at 58-69 %] because 16 of 32 lanes idle through
// a chain that is serial in tree depth and every matrix moves between lanes by shuffles.  For batches
// that fill the machine anyway the same arithmetic is done here by ONE THREAD per hand: the whole
// chain stays in registers, all global traffic is coalesced through hand-minor layouts (bone_t,
// dbone_t, dfeat_t written / read as 128-byte rows) or staged through a per-warp shared-memory tile,
// and a hand costs ~150 / ~300 warp-instructions.  It requires MANO's tree — a wrist with five
// chains of three joints (MB_MODEL_CHAINS_5X3) — which is what every MANO pickle has; other trees
// and small batches keep the warp-per-hand kernels.
#include <cuda_fp16.h>
#include "hand_math.cuh"
#include "blend_tc.cuh"
#include "ptx.cuh"

namespace mb {
namespace {

constexpr int LH_WARPS = 4;
constexpr int BP = 33;                               // staging pitch: element (row i, hand r) at i * 33 + r
constexpr int BUF_ROWS = FEAT_K;                     // 148 rows
constexpr int BUF_FLOATS = BUF_ROWS * BP;            // 4884 floats = 19.1 KB per warp
constexpr int THETA_ROW = 100;                       // theta[45] lives in rows 100..144 until its joint is processed
constexpr int PCA_PITCH = 48;

struct LhConsts {
    alignas(16) float pca[NAA * PCA_PITCH];          // [nc][48]
    float mean[NAA + 3];
    int pca_identity;                                // hands_components[:nc] is the identity (full axis-angle input): theta = mean + coeffs
    float j0[NJ * 3];
    alignas(16) float jb[NJ * 3 * 12];               // [48][12] (10 used)
};
// joints-only extras: the five fingertip vertices (333,444,672,555,745 -> joints 4,8,12,16,20) are evaluated from a
// 15-column slice of the blend basis and their dense skinning weights, without the 778-vertex contraction
constexpr int NTIP = 5;
__constant__ int c_tip_vert_lh[NTIP] = {333, 444, 672, 555, 745};
struct LhTips {
    alignas(16) float basis[FEAT_K][16];             // basis[f][tip*3 + c], column 15 = 0; row FEAT_ONE = v_template
    float w[NTIP][NJ];                               // dense skinning weights of the tip vertices
};

__device__ void lh_stage_constants(LhConsts& C, const void* blob, int nc) {
    const BlobLayout L = blob_layout();
    const float* pca = blob_ptr<float>(blob, L.pca);
    const float* mean = blob_ptr<float>(blob, L.pose_mean);
    const float* j0 = blob_ptr<float>(blob, L.j0);
    const float* jb = blob_ptr<float>(blob, L.jb);
    const int t = threadIdx.x, nt = blockDim.x;
    bool other = false;
    for (int i = t; i < NAA * PCA_PITCH; i += nt) {
        const int r = i / PCA_PITCH, c = i % PCA_PITCH;
        const float x = (r < nc && c < NAA) ? pca[r * NAA + c] : 0.f;
        C.pca[i] = x;
        other |= r < nc && x != (r == c ? 1.f : 0.f);
    }
    // "no PCA" models (BASELINE config 4: the 45 axis-angle values are the input) skip the 45 x 45 products;
    // block-uniform, decided from the constants themselves
    const int any_other = __syncthreads_or(other ? 1 : 0);
    if (t == 0) C.pca_identity = !any_other;
    for (int i = t; i < NAA; i += nt) C.mean[i] = mean[i];
    for (int i = t; i < NJ * 3; i += nt) C.j0[i] = j0[i];
    for (int i = t; i < NJ * 3 * 12; i += nt) {
        const int r = i / 12, c = i % 12;
        C.jb[i] = c < NB ? jb[r * NB + c] : 0.f;
    }
    __syncthreads();
}

__device__ void lh_stage_tips(LhTips& T, const void* blob) {
    const BlobLayout L = blob_layout();
    const float* basis = blob_ptr<float>(blob, L.basis);
    const float* sw = blob_ptr<float>(blob, L.skin_w);
    const uint8_t* sb = blob_ptr<uint8_t>(blob, L.skin_b);
    const int t = threadIdx.x, nt = blockDim.x;
    for (int i = t; i < FEAT_K * 16; i += nt) {
        const int k = i / 16, tc = i % 16;
        T.basis[k][tc] = tc < 15 ? basis[(size_t)k * VP_PITCH + c_tip_vert_lh[tc / 3] * 3 + tc % 3] : 0.f;
    }
    for (int i = t; i < NTIP * NJ; i += nt) {
        const int tp = i / NJ, b = i % NJ;
        float acc = 0.f;
        for (int sl = 0; sl < MAX_INFL; ++sl)
            if (sb[c_tip_vert_lh[tp] * MAX_INFL + sl] == b) acc += sw[c_tip_vert_lh[tp] * MAX_INFL + sl];
        T.w[tp][b] = acc;
    }
    __syncthreads();
}
// x * basis row f accumulated into the 15 tip coordinates
__device__ __forceinline__ void tip_axpy(const LhTips& T, int f, float x, float (&tv)[15]) {
    const float4* row = reinterpret_cast<const float4*>(T.basis[f]);
    const float4 a = row[0], b = row[1], c = row[2], d = row[3];
    tv[0] = fmaf(x, a.x, tv[0]); tv[1] = fmaf(x, a.y, tv[1]); tv[2] = fmaf(x, a.z, tv[2]); tv[3] = fmaf(x, a.w, tv[3]);
    tv[4] = fmaf(x, b.x, tv[4]); tv[5] = fmaf(x, b.y, tv[5]); tv[6] = fmaf(x, b.z, tv[6]); tv[7] = fmaf(x, b.w, tv[7]);
    tv[8] = fmaf(x, c.x, tv[8]); tv[9] = fmaf(x, c.y, tv[9]); tv[10] = fmaf(x, c.z, tv[10]); tv[11] = fmaf(x, c.w, tv[11]);
    tv[12] = fmaf(x, d.x, tv[12]); tv[13] = fmaf(x, d.y, tv[13]); tv[14] = fmaf(x, d.z, tv[14]);
}
// <basis row f, dtv> : the gradient of feature f that arrives through the tips
__device__ __forceinline__ float tip_dot(const LhTips& T, int f, const float (&dtv)[15]) {
    const float4* row = reinterpret_cast<const float4*>(T.basis[f]);
    const float4 a = row[0], b = row[1], c = row[2], d = row[3];
    float acc = a.x * dtv[0];
    acc = fmaf(a.y, dtv[1], acc); acc = fmaf(a.z, dtv[2], acc); acc = fmaf(a.w, dtv[3], acc);
    acc = fmaf(b.x, dtv[4], acc); acc = fmaf(b.y, dtv[5], acc); acc = fmaf(b.z, dtv[6], acc); acc = fmaf(b.w, dtv[7], acc);
    acc = fmaf(c.x, dtv[8], acc); acc = fmaf(c.y, dtv[9], acc); acc = fmaf(c.z, dtv[10], acc); acc = fmaf(c.w, dtv[11], acc);
    acc = fmaf(d.x, dtv[12], acc); acc = fmaf(d.y, dtv[13], acc); acc = fmaf(d.z, dtv[14], acc);
    return acc;
}
// rest-pose tip vertices: tv = v_template + S beta + P vec(R_j - I), theta read from staging rows th0..
__device__ __forceinline__ void tips_rest_pose(const LhTips& T, const float (&beta)[NB], const float* bl, int theta_row,
                                               float (&tv)[15]) {
#pragma unroll
    for (int c = 0; c < 15; ++c) tv[c] = T.basis[FEAT_ONE][c];
#pragma unroll
    for (int sft = 0; sft < NB; ++sft) tip_axpy(T, sft, beta[sft], tv);
#pragma unroll 1
    for (int k = 1; k < NJ; ++k) {
        const float* th = bl + (theta_row + 3 * (k - 1)) * BP;
        const M3 R = rodrigues(v3(th[0], th[BP], th[2 * BP]));
#pragma unroll
        for (int e = 0; e < 9; ++e) tip_axpy(T, NB + 9 * (k - 1) + e, R.m[e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f), tv);
    }
}

// coalesced copy of `n` consecutive floats (rows of up to 32 hands) into the warp's staging buffer
__device__ __forceinline__ void stage_in(float* dst, const float* __restrict__ src, int n, int lane) {
    warp_copy_async(dst, src, n, lane);                       // asynchronous; stage_wait() before the first read
}
__device__ __forceinline__ void stage_wait() {
    cp_async_wait_all();
    __syncwarp();
}

__device__ __forceinline__ V3 rest_joint(const LhConsts& C, int k, const float (&beta)[NB]) {
    float j[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* row = &C.jb[(k * 3 + c) * 12];
        float acc = C.j0[k * 3 + c];
#pragma unroll
        for (int s = 0; s < NB; ++s) acc = fmaf(row[s], beta[s], acc);
        j[c] = acc;
    }
    return v3(j[0], j[1], j[2]);
}

// theta = mean + coeffs . pca  -> rows THETA_ROW.. of the staging buffer (column = this lane)
__device__ __forceinline__ void lh_theta(const LhConsts& C, const float* s_coef, int nc, float* bl) {
    if (C.pca_identity) {
        for (int j = 0; j < NAA; ++j) bl[(THETA_ROW + j) * BP] = j < nc ? fmaf(s_coef[j], 1.f, C.mean[j]) : C.mean[j];
        return;
    }
    float th[NAA];
#pragma unroll
    for (int j = 0; j < NAA; ++j) th[j] = C.mean[j];
    for (int i = 0; i < nc; ++i) {
        const float ci = s_coef[i];
        const float4* row = reinterpret_cast<const float4*>(&C.pca[i * PCA_PITCH]);
#pragma unroll
        for (int q = 0; q < 11; ++q) {
            const float4 p = row[q];
            th[4 * q] = fmaf(ci, p.x, th[4 * q]); th[4 * q + 1] = fmaf(ci, p.y, th[4 * q + 1]);
            th[4 * q + 2] = fmaf(ci, p.z, th[4 * q + 2]); th[4 * q + 3] = fmaf(ci, p.w, th[4 * q + 3]);
        }
        th[44] = fmaf(ci, C.pca[i * PCA_PITCH + 44], th[44]);
    }
#pragma unroll
    for (int j = 0; j < NAA; ++j) bl[(THETA_ROW + j) * BP] = th[j];
}

// bone transform with the global rotation folded in: A' = [Rq Rg | Rq (tg - Rg J)]  -> bone_t[group][k][lane][12]
// (the forward chain runs in double and is rounded to fp32 here, once)
__device__ __forceinline__ void emit_bone(float* __restrict__ bt, int k, const M3d& Rq, const M3d& Rg, const V3d& tg, const V3d& J) {
    const M3 Rp = m3_from(m3d_mul(Rq, Rg));
    const V3 tp = v3_from(m3d_vec(Rq, v3d_sub(tg, m3d_vec(Rg, J))));
    float4* o = reinterpret_cast<float4*>(bt + k * (BONE_F * 32));      // bone_t[group][k][lane][12]
    o[0] = make_float4(Rp.m[0], Rp.m[1], Rp.m[2], tp.x);
    o[1] = make_float4(Rp.m[3], Rp.m[4], Rp.m[5], tp.y);
    o[2] = make_float4(Rp.m[6], Rp.m[7], Rp.m[8], tp.z);
}
__device__ __forceinline__ void emit_joint(float* __restrict__ jrow, int slot, const M3& Rq, const V3& tg) {
    const V3 j = m3_vec(Rq, tg);
    jrow[slot * 3] = j.x; jrow[slot * 3 + 1] = j.y; jrow[slot * 3 + 2] = j.z;
}
__device__ __forceinline__ void emit_joint(float* __restrict__ jrow, int slot, const M3d& Rq, const V3d& tg) {
    const V3 j = v3_from(m3d_vec(Rq, tg));
    jrow[slot * 3] = j.x; jrow[slot * 3 + 1] = j.y; jrow[slot * 3 + 2] = j.z;
}

__global__ void __launch_bounds__(LH_WARPS * 32)
pose_forward_lh_kernel(const void* __restrict__ blob, int nc, const float* __restrict__ rot,
                       const float* __restrict__ coeffs, const float* __restrict__ betas, int B,
                       float* __restrict__ feat, unsigned char* __restrict__ featp, float* __restrict__ bone_t,
                       float* __restrict__ joints) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LhConsts& C = *reinterpret_cast<LhConsts*>(smem_raw);
    lh_stage_constants(C, blob, nc);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* buf = reinterpret_cast<float*>(smem_raw + sizeof(LhConsts)) + warp * BUF_FLOATS;
    float* bl = buf + lane;                                    // (row i, this hand) at bl[i * BP]
    const int ngroups = (B + 31) >> 5;
    // block-uniform trip count with a barrier per pass: the four warps of a block walk this long straight-line code
    // (85-170 KB of SASS) together, so they share its instruction-cache lines instead of evicting each other's
    for (int g0 = blockIdx.x * LH_WARPS; g0 < ngroups; g0 += gridDim.x * LH_WARPS) {
        __syncthreads();
        const int g = g0 + warp;
        if (g >= ngroups) continue;
        const long long h0 = (long long)g * 32;
        const int nh = (B - h0) < 32 ? (int)(B - h0) : 32;
        const long long hand = h0 + (lane < nh ? lane : nh - 1);      // idle lanes of a ragged group redo the last hand
        const bool live = lane < nh;
        __syncwarp();
        // ---- inputs: coalesced rows -> staging (coeffs at [0, 32 nc), betas behind, rot behind)
        float* s_coef = buf;
        float* s_beta = buf + 32 * NAA;
        float* s_rot = s_beta + 32 * NB;
        stage_in(s_coef, coeffs + h0 * nc, nh * nc, lane);
        stage_in(s_beta, betas + h0 * NB, nh * NB, lane);
        stage_in(s_rot, rot + h0 * 3, nh * 3, lane);
        stage_wait();
        const int r = (int)(hand - h0);
        float beta[NB];
#pragma unroll
        for (int s = 0; s < NB; ++s) beta[s] = s_beta[r * NB + s];
        const M3d Rq = rodrigues_d(v3(s_rot[r * 3], s_rot[r * 3 + 1], s_rot[r * 3 + 2]));
        // theta needs this lane's coefficients: copy them out of the row-major staging first
        // (the theta rows alias nothing below float 32*NAA + 32*NB + 96 = 1856 < THETA_ROW * BP)
        lh_theta(C, s_coef + r * nc, nc, bl);
        __syncwarp();                                          // every lane is done with the input staging

        float* bt = bone_t + (size_t)g * (NJ * BONE_F * 32) + lane * BONE_F;
        float* jrow = joints + hand * (NOUTJ * 3);
        // ---- wrist: constant root rotation [pi, 0, 0] (:76, :128)
        const M3d R0 = rodrigues_d(v3(3.14159274101257324f, 0.f, 0.f));
        const V3d J0 = v3d(rest_joint(C, 0, beta));
        if (live) { emit_bone(bt, 0, Rq, R0, J0, J0); emit_joint(jrow, 0, Rq, J0); }
#pragma unroll
        for (int s = 0; s < NB; ++s) bl[s * BP] = beta[s];
        bl[FEAT_ONE * BP] = 1.f; bl[(FEAT_ONE + 1) * BP] = 0.f; bl[(FEAT_ONE + 2) * BP] = 0.f;
        // ---- five chains of three joints
#pragma unroll 1
        for (int f = 0; f < 5; ++f) {
            M3d Rgp = R0;
            V3d tgp = J0, Jp = J0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int k = 1 + 3 * f + i;
                const float* th = bl + (THETA_ROW + 3 * (k - 1)) * BP;
                const M3d R = rodrigues_d(v3(th[0], th[BP], th[2 * BP]));
                float* fr = bl + (NB + 9 * (k - 1)) * BP;
#pragma unroll
                for (int e = 0; e < 9; ++e) fr[e * BP] = (float)(R.m[e] - ((e == 0 || e == 4 || e == 8) ? 1.0 : 0.0));
                const V3d J = v3d(rest_joint(C, k, beta));
                const M3d Rg = m3d_mul(Rgp, R);
                const V3d tg = v3d_add(tgp, m3d_vec(Rgp, v3d_sub(J, Jp)));
                if (live) { emit_bone(bt, k, Rq, Rg, tg, J); emit_joint(jrow, 1 + 4 * f + i, Rq, tg); }
                Rgp = Rg; tgp = tg; Jp = J;
            }
        }
        // ---- blend features leave as fp32 rows (fp32 mode) or fp16 hi/lo UMMA K-groups (tensor-core modes)
        if (feat != nullptr) {
            __syncwarp();
            for (int rr = 0; rr < nh; ++rr) {
                float* o = feat + (h0 + rr) * FEAT_K;
                for (int i = lane; i < FEAT_K; i += 32) o[i] = buf[i * BP + rr];
            }
        }
        if (featp != nullptr && live) {
            const float fs = (float)(1 << TC_FEAT_SCALE_LOG2);
#pragma unroll 1
            for (int kg8 = 0; kg8 < TC_K / 8; ++kg8) {
                __align__(16) __half hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int k = kg8 * 8 + e;
                    const float x = k < TC_K_REAL ? bl[k * BP] * fs : 0.f;
                    hi[e] = __float2half_rn(x);
                    lo[e] = __float2half_rn(x - __half2float(hi[e]));
                }
                *reinterpret_cast<uint4*>(featp + tc_feat_group_offset(hand, kg8, 0)) = *reinterpret_cast<const uint4*>(hi);
                *reinterpret_cast<uint4*>(featp + tc_feat_group_offset(hand, kg8, 1)) = *reinterpret_cast<const uint4*>(lo);
            }
        }
    }
}

// ---- joints only (verts == NULL): the 21 joints without the 778-vertex contraction -------------------
__global__ void __launch_bounds__(LH_WARPS * 32)
pose_forward_lh_jo_kernel(const void* __restrict__ blob, int nc, const float* __restrict__ rot,
                          const float* __restrict__ coeffs, const float* __restrict__ betas, int B,
                          float* __restrict__ joints) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LhConsts& C = *reinterpret_cast<LhConsts*>(smem_raw);
    LhTips& T = *reinterpret_cast<LhTips*>(smem_raw + sizeof(LhConsts));
    lh_stage_constants(C, blob, nc);
    lh_stage_tips(T, blob);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* buf = reinterpret_cast<float*>(smem_raw + sizeof(LhConsts) + sizeof(LhTips)) + warp * BUF_FLOATS;
    float* bl = buf + lane;
    const int ngroups = (B + 31) >> 5;
    // block-uniform trip count with a barrier per pass: the four warps of a block walk this long straight-line code
    // (85-170 KB of SASS) together, so they share its instruction-cache lines instead of evicting each other's
    for (int g0 = blockIdx.x * LH_WARPS; g0 < ngroups; g0 += gridDim.x * LH_WARPS) {
        __syncthreads();
        const int g = g0 + warp;
        if (g >= ngroups) continue;
        const long long h0 = (long long)g * 32;
        const int nh = (B - h0) < 32 ? (int)(B - h0) : 32;
        const long long hand = h0 + (lane < nh ? lane : nh - 1);
        const bool live = lane < nh;
        __syncwarp();
        float* s_coef = buf;
        float* s_beta = buf + 32 * NAA;
        float* s_rot = s_beta + 32 * NB;
        stage_in(s_coef, coeffs + h0 * nc, nh * nc, lane);
        stage_in(s_beta, betas + h0 * NB, nh * NB, lane);
        stage_in(s_rot, rot + h0 * 3, nh * 3, lane);
        stage_wait();
        const int r = (int)(hand - h0);
        float beta[NB];
#pragma unroll
        for (int sft = 0; sft < NB; ++sft) beta[sft] = s_beta[r * NB + sft];
        const M3 Rq = rodrigues(v3(s_rot[r * 3], s_rot[r * 3 + 1], s_rot[r * 3 + 2]));
        lh_theta(C, s_coef + r * nc, nc, bl);
        __syncwarp();
        float tv[15];
        tips_rest_pose(T, beta, bl, THETA_ROW, tv);
        float out[15];
#pragma unroll
        for (int c = 0; c < 15; ++c) out[c] = 0.f;
        float* jrow = joints + hand * (NOUTJ * 3);
        // bone k is known: its joint, and its share of the five tips
        auto bone = [&](int k, int slot, const M3& Rg, const V3& tg, const V3& J) {
            const M3 Rp = m3_mul(Rq, Rg);
            const V3 tp = m3_vec(Rq, v3_sub(tg, m3_vec(Rg, J)));
            if (live) emit_joint(jrow, slot, Rq, tg);
#pragma unroll
            for (int t = 0; t < NTIP; ++t) {
                const float w = T.w[t][k];
                if (w != 0.f) {
                    const V3 p = m3_vec(Rp, v3(tv[3 * t], tv[3 * t + 1], tv[3 * t + 2]));
                    out[3 * t] = fmaf(w, p.x + tp.x, out[3 * t]);
                    out[3 * t + 1] = fmaf(w, p.y + tp.y, out[3 * t + 1]);
                    out[3 * t + 2] = fmaf(w, p.z + tp.z, out[3 * t + 2]);
                }
            }
        };
        const M3 R0 = rodrigues(v3(3.14159274101257324f, 0.f, 0.f));
        const V3 J0 = rest_joint(C, 0, beta);
        bone(0, 0, R0, J0, J0);
#pragma unroll 1
        for (int f = 0; f < 5; ++f) {
            M3 Rgp = R0;
            V3 tgp = J0, Jp = J0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int k = 1 + 3 * f + i;
                const float* th = bl + (THETA_ROW + 3 * (k - 1)) * BP;
                const M3 R = rodrigues(v3(th[0], th[BP], th[2 * BP]));
                const V3 J = rest_joint(C, k, beta);
                const M3 Rg = m3_mul(Rgp, R);
                const V3 tg = v3_add(tgp, m3_vec(Rgp, v3_sub(J, Jp)));
                bone(k, 1 + 4 * f + i, Rg, tg, J);
                Rgp = Rg; tgp = tg; Jp = J;
            }
        }
        if (live) {
#pragma unroll
            for (int t = 0; t < NTIP; ++t) {
                float* o = jrow + (4 + 4 * t) * 3;
                o[0] = out[3 * t]; o[1] = out[3 * t + 1]; o[2] = out[3 * t + 2];
            }
        }
    }
}

// ================================================================= backward
// Staging rows of the backward (per warp, pitch BP): dfeat rows 0..147 are loaded first; the chain
// joints' upstream gradients (48 floats) and the axis-angle gradients (45 floats) reuse rows that
// have been consumed.
// MODE 0: full backward (dfeat / dbone from the skinning and blend backward); MODE 1: joints only (the five tips stand
// in for the 778 vertices); MODE 2: one iteration of the fitting loop in ONE kernel (BASELINE config 5) — joints-only
// forward, gradient of the masked L2 objective against target keypoints (criterions/loss.py:10-25), joints-only
// backward, the regulariser's gradient (:113-117) and the Adam update of this hand's 58 parameters, plus this
// launch's partial sums {sum vis |d|^2, sum theta_new^2, sum beta_new^2}.  The batch-global quantities the gradient
// needs (visible count, Frobenius norms of the CURRENT parameters) depend only on the mask and the parameters, not on
// the forward pass, so they arrive reduced from the previous iteration (FitArgs::globals) and the iteration needs
// no grid-wide dependency inside.  In MODE 2 rot/coeffs/betas alias g_rot/g_coeffs/g_betas (updated in place).
struct FitArgs {
    const float* tgt;          // [B][21][3] target keypoints
    const float* vis;          // [B][21] visibility (non-zero = visible)
    float* m;                  // Adam first moments, laid out like the parameters: rot[B][3] | coeffs[B][nc] | betas[B][10]
    float* v;                  // Adam second moments
    const double* globals;     // device {N_vis, sum theta^2, sum beta^2} over ALL ranks for the current parameters
    double* partials;          // device {sum vis |d|^2, sum theta_new^2, sum beta_new^2} of this launch (zeroed by the launcher)
    float step_size, b1, b2, eps, inv_sqrt_bc2;
    int regularize;
};

// torch.optim.Adam on `n` consecutive parameters (rows of `w` values, gradients staged row-major with pitch `pitch`),
// plus `reg` * parameter added to the gradient; loads of a batch of ADAM_U rows are issued before any arithmetic so
// that one warp keeps 3 * ADAM_U coalesced requests in flight.  Returns this lane's sum of the updated squares.
constexpr int ADAM_U = 9;
__device__ __noinline__ float adam_rows(float* __restrict__ prm, float* __restrict__ m, float* __restrict__ v,
                                           const float* __restrict__ s_grad, int w, int pitch, int n, float reg,
                                           float b1, float b2, float step_size, float inv_sqrt_bc2, float eps, int lane) {
    const float omb1 = 1.f - b1, omb2 = 1.f - b2;
    float sq = 0.f;
    for (int base = lane; base < n; base += 32 * ADAM_U) {
        float pv[ADAM_U], mv[ADAM_U], vv[ADAM_U];
#pragma unroll
        for (int u = 0; u < ADAM_U; ++u) {
            const int i = base + 32 * u;
            if (i < n) { pv[u] = prm[i]; mv[u] = m[i]; vv[u] = v[i]; }
        }
#pragma unroll
        for (int u = 0; u < ADAM_U; ++u) {
            const int i = base + 32 * u;
            if (i < n) {
                const float gr = fmaf(reg, pv[u], s_grad[(i / w) * pitch + (i % w)]);
                const float mi = fmaf(b1, mv[u], omb1 * gr);
                const float vi = fmaf(b2, vv[u], omb2 * gr * gr);
                const float pn = pv[u] - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
                m[i] = mi; v[i] = vi; prm[i] = pn;
                sq = fmaf(pn, pn, sq);
            }
        }
    }
    return sq;
}

template <int MODE>
__global__ void __launch_bounds__(LH_WARPS * 32)
pose_backward_lh_kernel(const void* __restrict__ blob, int nc, const float* rot, const float* coeffs, const float* betas,
                        const float* __restrict__ dfeat_t, const float* __restrict__ dbone_t,
                        const float* __restrict__ g_joints, int B, float* g_rot, float* g_coeffs, float* g_betas,
                        const FitArgs F) {
    constexpr bool JO = MODE >= 1, FIT = MODE == 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LhConsts& C = *reinterpret_cast<LhConsts*>(smem_raw);
    LhTips& T = *reinterpret_cast<LhTips*>(smem_raw + sizeof(LhConsts));      // joints-only variant
    lh_stage_constants(C, blob, nc);
    if (JO) lh_stage_tips(T, blob);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* buf = reinterpret_cast<float*>(smem_raw + sizeof(LhConsts) + (JO ? sizeof(LhTips) : 0)) + warp * BUF_FLOATS;
    float* bl = buf + lane;
    const int ngroups = (B + 31) >> 5;
    // fitting step: batch-global scalars of the objective, this thread's partial sums
    float l2_scale = 0.f, reg_t = 0.f, reg_b = 0.f;
    double acc_s = 0.0, acc_t = 0.0, acc_b = 0.0;
    if (FIT) {
        const double nv = F.globals[0], tn = sqrt(F.globals[1]), bn = sqrt(F.globals[2]);
        l2_scale = nv > 0.0 ? (float)(2.0 / nv) : 0.f;                       // d(mean over visible of |d|^2) / d joint
        reg_t = (F.regularize && tn > 0.0) ? (float)(1.0 / (100.0 * tn)) : 0.f;   // d(|theta|_F / 100) = theta / (100 |theta|_F)
        reg_b = (F.regularize && bn > 0.0) ? (float)(1.0 / (10.0 * bn)) : 0.f;    // d(10 |beta|_F / 100)
    }
    // block-uniform trip count with a barrier per pass: the four warps of a block walk this long straight-line code
    // (85-170 KB of SASS) together, so they share its instruction-cache lines instead of evicting each other's
    for (int g0 = blockIdx.x * LH_WARPS; g0 < ngroups; g0 += gridDim.x * LH_WARPS) {
        __syncthreads();
        // a warp beyond the last group walks the pass with zero hands (it reads group ngroups-1, writes nothing): the
        // barriers inside the finger loops below need every warp of the block
        const bool active = g0 + warp < ngroups;
        const int g = active ? g0 + warp : ngroups - 1;
        const long long h0 = (long long)g * 32;
        const int nh = !active ? 0 : ((B - h0) < 32 ? (int)(B - h0) : 32);
        const long long hand = h0 + (lane < nh ? lane : (nh > 0 ? nh - 1 : 0));
        const bool live = lane < nh;
        const int r = (int)(hand - h0);
        __syncwarp();
        float* s_coef = buf;
        float* s_beta = buf + 32 * NAA;
        float* s_rot = s_beta + 32 * NB;
        float* s_gj = s_rot + 96;                              // [32][63] upstream joint gradients (fit: targets), row-major
        float* s_vis = s_gj + 32 * NOUTJ * 3;                  // [32][21] fit only: 3872 + 672 = 4544 floats < BUF_FLOATS
        stage_in(s_coef, coeffs + h0 * nc, nh * nc, lane);
        stage_in(s_beta, betas + h0 * NB, nh * NB, lane);
        stage_in(s_rot, rot + h0 * 3, nh * 3, lane);
        stage_in(s_gj, (FIT ? F.tgt : g_joints) + h0 * (NOUTJ * 3), nh * NOUTJ * 3, lane);   // 1856 + 2016 = 3872 floats < BUF_FLOATS
        if (FIT) stage_in(s_vis, F.vis + h0 * NOUTJ, nh * NOUTJ, lane);
        stage_wait();
        float beta[NB];
#pragma unroll
        for (int s = 0; s < NB; ++s) beta[s] = s_beta[r * NB + s];
        const V3 rq = v3(s_rot[r * 3], s_rot[r * 3 + 1], s_rot[r * 3 + 2]);
        const M3 Rq = rodrigues(rq);
        // upstream gradients of the 16 chain joints -> registers (slot of chain joint k: 0, 1+4f+i)
        V3 gj[NJ];
        float gt[15];                                          // upstream gradients of the five tip joints (joints-only)
        auto load_joint_grads = [&]() {
            gj[0] = v3(s_gj[r * 63], s_gj[r * 63 + 1], s_gj[r * 63 + 2]);
#pragma unroll
            for (int f = 0; f < 5; ++f)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float* p = s_gj + r * 63 + (1 + 4 * f + i) * 3;
                    gj[1 + 3 * f + i] = v3(p[0], p[1], p[2]);
                }
#pragma unroll
            for (int t = 0; t < NTIP; ++t)
#pragma unroll
                for (int c = 0; c < 3; ++c) gt[3 * t + c] = JO ? s_gj[r * 63 + (4 + 4 * t) * 3 + c] : 0.f;
        };
        if (!FIT) load_joint_grads();
        float th[NAA];
        {
            // theta in registers (the staging buffer is about to be overwritten by dfeat)
#pragma unroll
            for (int j = 0; j < NAA; ++j) th[j] = C.mean[j];
            const float* sc = s_coef + r * nc;
            if (C.pca_identity) {
#pragma unroll
                for (int j = 0; j < NAA; ++j) if (j < nc) th[j] = fmaf(sc[j], 1.f, th[j]);
            } else
            for (int i = 0; i < nc; ++i) {
                const float ci = sc[i];
                const float4* row = reinterpret_cast<const float4*>(&C.pca[i * PCA_PITCH]);
#pragma unroll
                for (int q = 0; q < 11; ++q) {
                    const float4 p = row[q];
                    th[4 * q] = fmaf(ci, p.x, th[4 * q]); th[4 * q + 1] = fmaf(ci, p.y, th[4 * q + 1]);
                    th[4 * q + 2] = fmaf(ci, p.z, th[4 * q + 2]); th[4 * q + 3] = fmaf(ci, p.w, th[4 * q + 3]);
                }
                th[44] = fmaf(ci, C.pca[i * PCA_PITCH + 44], th[44]);
            }
        }
        __syncwarp();                                          // input staging consumed by every lane
        // theta and the joint gradients move to staging rows (dynamic joint index inside the chain loop);
        // dfeat is read straight from its hand-minor global rows
#pragma unroll
        for (int j = 0; j < NAA; ++j) bl[j * BP] = th[j];                      // rows 0..44: theta, later dtheta
        float tv[15], dtv[15];
#pragma unroll
        for (int c = 0; c < 15; ++c) { tv[c] = 0.f; dtv[c] = 0.f; }
        if (FIT) {
            // ---- forward (as pose_forward_lh_jo_kernel): every joint turns its staged target into the objective's
            // gradient in place — row r of s_gj belongs to this lane alone (idle lanes of a ragged group do not write)
            tips_rest_pose(T, beta, bl, 0, tv);
            float sres = 0.f;
            auto resid = [&](int slot, const V3& j) {
                float* p = s_gj + r * 63 + slot * 3;
                const float dx = j.x - p[0], dy = j.y - p[1], dz = j.z - p[2];
                const bool on = s_vis[r * NOUTJ + slot] != 0.f;
                if (on) sres += dx * dx + dy * dy + dz * dz;
                const float sc = on ? l2_scale : 0.f;
                if (live) { p[0] = sc * dx; p[1] = sc * dy; p[2] = sc * dz; }
            };
            float out[15];
#pragma unroll
            for (int c = 0; c < 15; ++c) out[c] = 0.f;
            auto bone = [&](int k, int slot, const M3& Rg, const V3& tg, const V3& J) {
                const M3 Rp = m3_mul(Rq, Rg);
                const V3 tp = m3_vec(Rq, v3_sub(tg, m3_vec(Rg, J)));
                resid(slot, m3_vec(Rq, tg));
#pragma unroll
                for (int t = 0; t < NTIP; ++t) {
                    const float w = T.w[t][k];
                    if (w != 0.f) {
                        const V3 p = m3_vec(Rp, v3(tv[3 * t], tv[3 * t + 1], tv[3 * t + 2]));
                        out[3 * t] = fmaf(w, p.x + tp.x, out[3 * t]);
                        out[3 * t + 1] = fmaf(w, p.y + tp.y, out[3 * t + 1]);
                        out[3 * t + 2] = fmaf(w, p.z + tp.z, out[3 * t + 2]);
                    }
                }
            };
            const M3 R0f = rodrigues(v3(3.14159274101257324f, 0.f, 0.f));
            const V3 J0f = rest_joint(C, 0, beta);
            bone(0, 0, R0f, J0f, J0f);
#pragma unroll 1
            for (int f = 0; f < 5; ++f) {
                __syncthreads();                               // keep the block's warps on the same code
                M3 Rgp = R0f;
                V3 tgp = J0f, Jp = J0f;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int k = 1 + 3 * f + i;
                    const float* tp = bl + (3 * (k - 1)) * BP;
                    const M3 R = rodrigues(v3(tp[0], tp[BP], tp[2 * BP]));
                    const V3 J = rest_joint(C, k, beta);
                    const M3 Rg = m3_mul(Rgp, R);
                    const V3 tg = v3_add(tgp, m3_vec(Rgp, v3_sub(J, Jp)));
                    bone(k, 1 + 4 * f + i, Rg, tg, J);
                    Rgp = Rg; tgp = tg; Jp = J;
                }
            }
#pragma unroll
            for (int t = 0; t < NTIP; ++t) resid(4 + 4 * t, v3(out[3 * t], out[3 * t + 1], out[3 * t + 2]));
            if (live) acc_s += (double)sres;
            load_joint_grads();
            __syncwarp();                                      // targets / visibility consumed: rows 48.. may be overwritten
        }
#pragma unroll
        for (int k = 0; k < NJ; ++k) { bl[(48 + 3 * k) * BP] = gj[k].x; bl[(49 + 3 * k) * BP] = gj[k].y; bl[(50 + 3 * k) * BP] = gj[k].z; }
        const float* df = JO ? nullptr : dfeat_t + (size_t)g * (TC_K * 32) + lane;             // dfeat_t[g][k][lane], k < 160
        const float* db = JO ? nullptr : dbone_t + (size_t)g * (NJ * BONE_F * 32) + lane;      // dbone_t[g][bone*12+e][lane]

        // joints only: the five tips stand in for the 778 vertices — rest-pose tips tv, and (pre-pass over
        // the chain) d tv = sum_k w_tk R'_k^T g_t, from which every feature gradient follows
        if (JO) {
            if (!FIT) tips_rest_pose(T, beta, bl, 0, tv);
            auto tip_back = [&](int k, const M3& Rg) {
                const M3 Rp = m3_mul(Rq, Rg);
#pragma unroll
                for (int t = 0; t < NTIP; ++t) {
                    const float w = T.w[t][k];
                    if (w != 0.f) {
                        const V3 d = m3_tvec(Rp, v3(gt[3 * t], gt[3 * t + 1], gt[3 * t + 2]));
                        dtv[3 * t] = fmaf(w, d.x, dtv[3 * t]); dtv[3 * t + 1] = fmaf(w, d.y, dtv[3 * t + 1]); dtv[3 * t + 2] = fmaf(w, d.z, dtv[3 * t + 2]);
                    }
                }
            };
            const M3 R0p = rodrigues(v3(3.14159274101257324f, 0.f, 0.f));
            tip_back(0, R0p);
#pragma unroll 1
            for (int f = 0; f < 5; ++f) {
                __syncthreads();                               // keep the block's warps on the same code
                M3 Rgp = R0p;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int k = 1 + 3 * f + i;
                    const float* tp = bl + (3 * (k - 1)) * BP;
                    const M3 Rg = m3_mul(Rgp, rodrigues(v3(tp[0], tp[BP], tp[2 * BP])));
                    tip_back(k, Rg);
                    Rgp = Rg;
                }
            }
        }
        float gbeta[NB];
#pragma unroll
        for (int s = 0; s < NB; ++s) gbeta[s] = JO ? tip_dot(T, s, dtv) : df[s * 32];
        M3 dRq = m3_zero();
        // d beta += Jb^T dJ for one joint
        auto add_dJ = [&](int k, const V3& dJ) {
            const float d3[3] = {dJ.x, dJ.y, dJ.z};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* row = &C.jb[(k * 3 + c) * 12];
#pragma unroll
                for (int s = 0; s < NB; ++s) gbeta[s] = fmaf(row[s], d3[c], gbeta[s]);
            }
        };
        // A.2 step 3 for one bone: split dA'_k = d[Rq Rg | Rq tA], joint_k = Rq tg
        auto split_bone = [&](int k, const M3& Rg, const V3& tg, const V3& J, const V3& gjk, M3& dRg, V3& dtg, V3& dJ) {
            float dA[BONE_F];
            if (JO) {
                // dA'_k = sum_t w_tk g_t (x) [tv_t ; 1]
#pragma unroll
                for (int e = 0; e < BONE_F; ++e) dA[e] = 0.f;
#pragma unroll
                for (int t = 0; t < NTIP; ++t) {
                    const float w = T.w[t][k];
                    if (w != 0.f) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const float wg = w * gt[3 * t + i];
                            dA[4 * i] = fmaf(wg, tv[3 * t], dA[4 * i]); dA[4 * i + 1] = fmaf(wg, tv[3 * t + 1], dA[4 * i + 1]);
                            dA[4 * i + 2] = fmaf(wg, tv[3 * t + 2], dA[4 * i + 2]); dA[4 * i + 3] += wg;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < BONE_F; ++e) dA[e] = db[(k * BONE_F + e) * 32];
            }
            const V3 tA = v3_sub(tg, m3_vec(Rg, J));
            M3 dApR;
            dApR.m[0] = dA[0]; dApR.m[1] = dA[1]; dApR.m[2] = dA[2];
            dApR.m[3] = dA[4]; dApR.m[4] = dA[5]; dApR.m[5] = dA[6];
            dApR.m[6] = dA[8]; dApR.m[7] = dA[9]; dApR.m[8] = dA[10];
            const V3 dApt = v3(dA[3], dA[7], dA[11]);
            // dRq += dA'R Rg^T + dA't tA^T + gj tg^T
            m3_acc(dRq, m3_mult(dApR, Rg));
            m3_add_outer(dRq, dApt, tA);
            m3_add_outer(dRq, gjk, tg);
            dRg = m3_tmul(Rq, dApR);
            const V3 dAt = m3_tvec(Rq, dApt);
            m3_add_outer(dRg, v3(-dAt.x, -dAt.y, -dAt.z), J);
            dtg = v3_add(dAt, m3_tvec(Rq, gjk));
            dJ = m3_tvec(Rg, v3(-dAt.x, -dAt.y, -dAt.z));
        };

        const M3 R0 = rodrigues(v3(3.14159274101257324f, 0.f, 0.f));
        const V3 J0 = rest_joint(C, 0, beta);
        M3 dRg0; V3 dtg0, dJ0;
        split_bone(0, R0, J0, J0, v3(bl[48 * BP], bl[49 * BP], bl[50 * BP]), dRg0, dtg0, dJ0);
#pragma unroll 1
        for (int f = 0; f < 5; ++f) {
            __syncthreads();                                   // keep the block's warps on the same code
            // forward through the chain (state of the three joints stays in registers)
            M3 R[3], Rg[3];
            V3 tg[3], J[3], rr[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int k = 1 + 3 * f + i;
                const float* tp = bl + (3 * (k - 1)) * BP;
                rr[i] = v3(tp[0], tp[BP], tp[2 * BP]);
                R[i] = rodrigues(rr[i]);
                J[i] = rest_joint(C, k, beta);
                const M3& Rgp = i == 0 ? R0 : Rg[i - 1];
                const V3& tgp = i == 0 ? J0 : tg[i - 1];
                const V3& Jp = i == 0 ? J0 : J[i - 1];
                Rg[i] = m3_mul(Rgp, R[i]);
                tg[i] = v3_add(tgp, m3_vec(Rgp, v3_sub(J[i], Jp)));
            }
            // reverse: tip of the chain first
            M3 dRg_c = m3_zero();                              // gradient arriving from the child
            V3 dtg_c = v3(0.f, 0.f, 0.f), dJ_c = v3(0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 2; i >= 0; --i) {
                const int k = 1 + 3 * f + i;
                const float* gp = bl + (48 + 3 * k) * BP;
                M3 dRg; V3 dtg, dJ;
                split_bone(k, Rg[i], tg[i], J[i], v3(gp[0], gp[BP], gp[2 * BP]), dRg, dtg, dJ);
                m3_acc(dRg, dRg_c); dtg = v3_add(dtg, dtg_c); dJ = v3_add(dJ, dJ_c);
                const M3& Rgp = i == 0 ? R0 : Rg[i - 1];
                const V3& Jp = i == 0 ? J0 : J[i - 1];
                // A.2 step 4: Rg = Rgp R ; tg = tgp + Rgp (J - Jp)
                M3 dRl = m3_tmul(Rgp, dRg);
                M3 up = m3_mult(dRg, R[i]);
                m3_add_outer(up, dtg, v3_sub(J[i], Jp));
                const V3 dd = m3_tvec(Rgp, dtg);
                dJ = v3_add(dJ, dd);
                add_dJ(k, dJ);
                // A.2 steps 5-6: pose-feature gradient joins dR_k ; Rodrigues backward
#pragma unroll
                for (int e = 0; e < 9; ++e) dRl.m[e] += JO ? tip_dot(T, NB + 9 * (k - 1) + e, dtv) : df[(NB + 9 * (k - 1) + e) * 32];
                const V3 dth = rodrigues_bwd(rr[i], dRl);
                float* tp = bl + (3 * (k - 1)) * BP;
                tp[0] = dth.x; tp[BP] = dth.y; tp[2 * BP] = dth.z;     // theta row -> dtheta row
                dRg_c = up; dtg_c = dtg; dJ_c = v3(-dd.x, -dd.y, -dd.z);
            }
            m3_acc(dRg0, dRg_c); dtg0 = v3_add(dtg0, dtg_c); dJ0 = v3_add(dJ0, dJ_c);
        }
        dJ0 = v3_add(dJ0, dtg0);                               // tg_0 = J_0 ; R_0 is a constant
        add_dJ(0, dJ0);
        const V3 drq = rodrigues_bwd(rq, dRq);

        // ---- A.2 step 7: d_coeffs = C[:nc] dtheta
        float dth[NAA];
#pragma unroll
        for (int j = 0; j < NAA; ++j) dth[j] = bl[j * BP];
        __syncwarp();
        float* s_gc = buf + 64 * BP;                           // [32][nc | 1] row-major, rows 64.. (theta rows are consumed)
        const int gp = nc | 1;
        if (C.pca_identity) {
#pragma unroll
            for (int j = 0; j < NAA; ++j) if (j < nc) s_gc[lane * gp + j] = dth[j];
        } else
        for (int i = 0; i < nc; ++i) {
            const float4* row = reinterpret_cast<const float4*>(&C.pca[i * PCA_PITCH]);
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < 11; ++q) {
                const float4 p = row[q];
                acc = fmaf(p.x, dth[4 * q], acc); acc = fmaf(p.y, dth[4 * q + 1], acc);
                acc = fmaf(p.z, dth[4 * q + 2], acc); acc = fmaf(p.w, dth[4 * q + 3], acc);
            }
            acc = fmaf(C.pca[i * PCA_PITCH + 44], dth[44], acc);
            s_gc[lane * gp + i] = acc;
        }
        __syncwarp();
        if (!FIT) {
            for (int i = lane; i < nh * nc; i += 32) g_coeffs[h0 * nc + i] = s_gc[(i / nc) * gp + (i % nc)];
            if (live) {
                g_rot[hand * 3] = drq.x; g_rot[hand * 3 + 1] = drq.y; g_rot[hand * 3 + 2] = drq.z;
#pragma unroll
                for (int s = 0; s < NB; ++s) g_betas[hand * NB + s] = gbeta[s];
            }
        } else {
            // ---- regulariser gradient + torch.optim.Adam (as adam_kernel in reduce.cu), parameters updated in place
            const size_t off_c = (size_t)B * 3, off_b = (size_t)B * (3 + nc);
            // rot / beta gradients join the coefficient gradients in the row-major staging so that every parameter,
            // moment and update moves as a coalesced row of the flat buffers
            float* s_gb = s_gc + 32 * 46;                      // [32][11]
            float* s_gr = s_gb + 32 * 11;                      // [32][3]   (ends at float 2112 + 1472 + 352 + 96 < BUF_FLOATS)
#pragma unroll
            for (int s = 0; s < NB; ++s) s_gb[lane * 11 + s] = fmaf(reg_b, beta[s], gbeta[s]);
            s_gr[lane * 3] = drq.x; s_gr[lane * 3 + 1] = drq.y; s_gr[lane * 3 + 2] = drq.z;
            __syncwarp();
            const float st = adam_rows(g_coeffs + (size_t)h0 * nc, F.m + off_c + (size_t)h0 * nc, F.v + off_c + (size_t)h0 * nc,
                                       s_gc, nc, gp, nh * nc, reg_t, F.b1, F.b2, F.step_size, F.inv_sqrt_bc2, F.eps, lane);
            const float sb = adam_rows(g_betas + (size_t)h0 * NB, F.m + off_b + (size_t)h0 * NB, F.v + off_b + (size_t)h0 * NB,
                                       s_gb, NB, 11, nh * NB, 0.f, F.b1, F.b2, F.step_size, F.inv_sqrt_bc2, F.eps, lane);
            adam_rows(g_rot + (size_t)h0 * 3, F.m + (size_t)h0 * 3, F.v + (size_t)h0 * 3, s_gr, 3, 3, nh * 3, 0.f, F.b1, F.b2, F.step_size, F.inv_sqrt_bc2, F.eps, lane);
            acc_t += (double)st;
            acc_b += (double)sb;
        }
    }
    if (FIT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc_s += __shfl_xor_sync(0xffffffffu, acc_s, o);
            acc_t += __shfl_xor_sync(0xffffffffu, acc_t, o);
            acc_b += __shfl_xor_sync(0xffffffffu, acc_b, o);
        }
        if (lane == 0) {
            atomicAdd(&F.partials[0], acc_s); atomicAdd(&F.partials[1], acc_t); atomicAdd(&F.partials[2], acc_b);
        }
    }
}

constexpr size_t LH_SMEM = sizeof(LhConsts) + (size_t)LH_WARPS * BUF_FLOATS * sizeof(float);
constexpr size_t LH_SMEM_JO = LH_SMEM + sizeof(LhTips);

inline int lh_grid(int B) {
    const long long groups = ((long long)B + 31) / 32;
    const long long blocks = (groups + LH_WARPS - 1) / LH_WARPS;
    const long long cap = (long long)NUM_SMS * (8 / LH_WARPS);
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace

int launch_pose_forward_lh(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                           int B, float* feat, unsigned char* featp, float* bone_t, float* joints, cudaStream_t s) {
    static SmemAttrOnce once;
    if (int arc = ensure_dyn_smem(once, pose_forward_lh_kernel, LH_SMEM)) return arc;
    pose_forward_lh_kernel<<<lh_grid(B), LH_WARPS * 32, LH_SMEM, s>>>(blob, nc, rot, coeffs, betas, B, feat, featp, bone_t, joints);
    return cuda_rc();
}

int launch_pose_backward_lh(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                            const float* dfeat_t, const float* dbone_t, const float* g_joints, int B,
                            float* g_rot, float* g_coeffs, float* g_betas, cudaStream_t s) {
    static SmemAttrOnce once;
    if (int arc = ensure_dyn_smem(once, pose_backward_lh_kernel<0>, LH_SMEM)) return arc;
    pose_backward_lh_kernel<0><<<lh_grid(B), LH_WARPS * 32, LH_SMEM, s>>>(blob, nc, rot, coeffs, betas, dfeat_t, dbone_t, g_joints, B,
                                                                             g_rot, g_coeffs, g_betas, FitArgs{});
    return cuda_rc();
}

int launch_joints_only_forward_lh(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                                  int B, float* joints, cudaStream_t s) {
    static SmemAttrOnce once;
    if (int arc = ensure_dyn_smem(once, pose_forward_lh_jo_kernel, LH_SMEM_JO)) return arc;
    pose_forward_lh_jo_kernel<<<lh_grid(B), LH_WARPS * 32, LH_SMEM_JO, s>>>(blob, nc, rot, coeffs, betas, B, joints);
    return cuda_rc();
}

int launch_joints_only_backward_lh(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                                   const float* g_joints, int B, float* g_rot, float* g_coeffs, float* g_betas, cudaStream_t s) {
    static SmemAttrOnce once;
    if (int arc = ensure_dyn_smem(once, pose_backward_lh_kernel<1>, LH_SMEM_JO)) return arc;
    pose_backward_lh_kernel<1><<<lh_grid(B), LH_WARPS * 32, LH_SMEM_JO, s>>>(blob, nc, rot, coeffs, betas, nullptr, nullptr, g_joints, B,
                                                                               g_rot, g_coeffs, g_betas, FitArgs{});
    return cuda_rc();
}

// One fitting iteration (MODE 2 above).  params / exp_avg / exp_avg_sq: rot[B][3] | coeffs[B][nc] | betas[B][10].
int launch_fit_step_lh(const void* blob, int nc, float* params, float* exp_avg, float* exp_avg_sq, const float* target_joints,
                       const float* vis, int B, const double* globals, double* partials, float lr, float beta1, float beta2,
                       float eps, int step, int regularize, cudaStream_t s) {
    static SmemAttrOnce once;
    if (int arc = ensure_dyn_smem(once, pose_backward_lh_kernel<2>, LH_SMEM_JO)) return arc;
    cudaError_t e = cudaMemsetAsync(partials, 0, 3 * sizeof(double), s);
    if (e != cudaSuccess) return (int)e;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    FitArgs F;
    F.tgt = target_joints; F.vis = vis; F.m = exp_avg; F.v = exp_avg_sq; F.globals = globals; F.partials = partials;
    F.step_size = (float)(lr / bc1); F.b1 = beta1; F.b2 = beta2; F.eps = eps; F.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    F.regularize = regularize;
    float* rot = params;
    float* coeffs = params + (size_t)B * 3;
    float* betas = params + (size_t)B * (3 + nc);
    pose_backward_lh_kernel<2><<<lh_grid(B), LH_WARPS * 32, LH_SMEM_JO, s>>>(blob, nc, rot, coeffs, betas, nullptr, nullptr, nullptr, B,
                                                                            rot, coeffs, betas, F);
    return cuda_rc();
}

}  // namespace mb
