// api.cu — extern "C" entry points of the MANO path (see include/mano_b200.h).
#include <string.h>
#include "common.cuh"

using namespace mb;

#include "blend_tc.cuh"
#include "skin.cuh"
#include <stdlib.h>
#include "vskin.cuh"

// ---------------------------------------------------------------- bookkeeping
#include <atomic>
#include <mutex>
#include <vector>
#include <algorithm>
namespace mb {
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfRec { int stage; cudaEvent_t a, b; };
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec*> g_prof_recs;

StageTimer::StageTimer(int st, cudaStream_t s) : stage(st), stream(s), rec(nullptr) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfRec* r = new ProfRec;
    r->stage = st;
    if (cudaEventCreate(&r->a) != cudaSuccess || cudaEventCreate(&r->b) != cudaSuccess) { delete r; return; }
    cudaEventRecord(r->a, s);
    rec = r;
}
StageTimer::~StageTimer() {
    if (!rec) return;
    ProfRec* r = reinterpret_cast<ProfRec*>(rec);
    cudaEventRecord(r->b, stream);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back(r);
}
}  // namespace mb

extern "C" long long mb_launch_count(void) { return g_launches.load(); }
extern "C" void mb_profile_enable(int on) { g_prof_on.store(on ? 1 : 0); }
extern "C" int mb_profile_collect(double* ms, long long* counts) {
    if (!ms || !counts) return MB_E_NULL;
    for (int i = 0; i < ST_COUNT; ++i) { ms[i] = 0.0; counts[i] = 0; }
    std::lock_guard<std::mutex> lk(g_prof_mu);
    int rc = 0;
    for (ProfRec* r : g_prof_recs) {
        cudaError_t e = cudaEventSynchronize(r->b);
        float t = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t, r->a, r->b);
        if (e == cudaSuccess) { ms[r->stage] += t; counts[r->stage] += 1; } else rc = (int)e;
        cudaEventDestroy(r->a); cudaEventDestroy(r->b);
        delete r;
    }
    g_prof_recs.clear();
    return rc;
}

extern "C" int mb_abi_version(void) { return MB_ABI_VERSION; }

extern "C" const char* mb_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case MB_E_NULL: return "mano_b200: a required pointer is NULL";
        case MB_E_RANGE: return "mano_b200: argument out of range (B < 0, pose_num outside [1,45], unknown mode/kind)";
        case MB_E_WORKSPACE: return "mano_b200: workspace too small (see mb_mano_workspace_bytes)";
        case MB_E_ALIGN: return "mano_b200: pointer must be 16-byte aligned";
        case MB_E_MODEL: return "mano_b200: model outside supported limits (<= 8 bones per vertex, <= 3200 skin weights, parents before children)";
        case MB_E_DEVICE: return "mano_b200: this library only runs on compute capability 10.x (B200, sm_100a)";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "mano_b200: unknown error";
}

extern "C" int mb_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return (int)e;
    return major == 10 ? 0 : MB_E_DEVICE;
}

extern "C" size_t mb_mano_blob_bytes(void) { return blob_layout().total + blend_tc_blob_bytes() + vskin_blob_bytes(); }

extern "C" int mb_mano_pack_constants(const float* basis, const float* j0, const float* jb, const float* pca, int nc,
                                      const float* pose_mean, const float* skin_w, const int32_t* skin_b,
                                      const int32_t* parents, void* host_blob) {
    if (!basis || !j0 || !jb || !pca || !pose_mean || !skin_w || !skin_b || !parents || !host_blob) return MB_E_NULL;
    if (nc < 1 || nc > NAA) return MB_E_RANGE;
    const BlobLayout L = blob_layout();
    char* out = reinterpret_cast<char*>(host_blob);
    memset(out, 0, L.total);

    BlobHeader* H = reinterpret_cast<BlobHeader*>(out + L.header);
    H->magic = 0x4d423230;
    H->abi = MB_ABI_VERSION;
    H->nc = nc;
    if (parents[0] >= 0) return MB_E_MODEL;
    int maxd = 0;
    for (int i = 0; i < NJ; ++i) {
        H->parents[i] = parents[i];
        if (i > 0 && (parents[i] < 0 || parents[i] >= i)) return MB_E_MODEL;
        H->depth[i] = i == 0 ? 0 : H->depth[parents[i]] + 1;
        if (H->depth[i] > maxd) maxd = H->depth[i];
    }
    H->max_depth = maxd;
    for (int i = 1; i < NJ; ++i) {
        const int p = parents[i];
        H->children[p][H->n_children[p]++] = i;
    }
    for (int i = 0; i < NJ; ++i)
        if (H->n_children[i] > H->max_children_at_depth[H->depth[i]]) H->max_children_at_depth[H->depth[i]] = H->n_children[i];

    float* b = reinterpret_cast<float*>(out + L.basis);
    float* bt = reinterpret_cast<float*>(out + L.basis_t);
    for (int k = 0; k < FEAT_K; ++k)
        for (int c = 0; c < NVC; ++c) {
            const float v = basis[(size_t)k * NVC + c];
            b[(size_t)k * VP_PITCH + c] = v;
            bt[(size_t)c * FEAT_K + k] = v;
        }
    memcpy(out + L.j0, j0, sizeof(float) * NJ * 3);
    memcpy(out + L.jb, jb, sizeof(float) * NJ * 3 * NB);
    memcpy(out + L.pca, pca, sizeof(float) * nc * NAA);
    memcpy(out + L.pose_mean, pose_mean, sizeof(float) * NAA);

    float* sw = reinterpret_cast<float*>(out + L.skin_w);
    uint8_t* sb = reinterpret_cast<uint8_t*>(out + L.skin_b);
    uint8_t* sc = reinterpret_cast<uint8_t*>(out + L.skin_cnt);
    int* cptr = reinterpret_cast<int*>(out + L.csc_ptr);
    int* cv = reinterpret_cast<int*>(out + L.csc_v);
    float* cw = reinterpret_cast<float*>(out + L.csc_w);
    int per_bone[NJ] = {0};
    int nnz = 0;
    for (int v = 0; v < NV; ++v) {
        int cnt = 0;
        for (int s = 0; s < MAX_INFL; ++s) {
            const float w = skin_w[v * MAX_INFL + s];
            const int bi = skin_b[v * MAX_INFL + s];
            if (w == 0.f) continue;
            if (bi < 0 || bi >= NJ) return MB_E_MODEL;
            sw[v * MAX_INFL + cnt] = w;
            sb[v * MAX_INFL + cnt] = (uint8_t)bi;
            ++cnt; ++per_bone[bi]; ++nnz;
        }
        sc[v] = (uint8_t)cnt;
    }
    if (nnz > MAX_NNZ) return MB_E_MODEL;
    cptr[0] = 0;
    for (int k = 0; k < NJ; ++k) cptr[k + 1] = cptr[k] + per_bone[k];
    int fill[NJ];
    for (int k = 0; k < NJ; ++k) fill[k] = cptr[k];
    for (int v = 0; v < NV; ++v)
        for (int s = 0; s < sc[v]; ++s) {
            const int bi = sb[v * MAX_INFL + s];
            cv[fill[bi]] = v;
            cw[fill[bi]] = sw[v * MAX_INFL + s];
            ++fill[bi];
        }
    H->csc_nnz = nnz;

    // ---- skin program of the register-blocked skinning kernels + block-order v_template ----
    std::vector<int32_t> coord_map(SK_TMPL_PAD);
    {
        const int rc = skin_pack(skin_w, skin_b, host_blob, coord_map.data());
        if (rc) return rc;
        float* tm = reinterpret_cast<float*>(out + L.sk_tmpl);
        for (int c = 0; c < SK_TMPL_PAD; ++c) tm[c] = coord_map[c] >= 0 ? basis[(size_t)FEAT_ONE * NVC + coord_map[c]] : 0.f;
    }
    blend_tc_pack(basis, coord_map.data(), out + L.total);
    // operand images of the fused lane = vertex forward (vskin.cu): same power-of-two basis scale as the blend images
    vskin_pack(basis, skin_w, skin_b, reinterpret_cast<const int32_t*>(out + L.sk_perm),
               reinterpret_cast<const TcBlobHeader*>(out + L.total)->basis_scale_log2, out + L.total + blend_tc_blob_bytes());
    return skin_program_check(host_blob, nullptr);
}

extern "C" int mb_mano_skin_program_stats(const void* host_blob, int32_t* stats4) {
    if (!host_blob || !stats4) return MB_E_NULL;
    return skin_program_check(host_blob, stats4);
}

extern "C" int mb_mano_model_flags(const int32_t* parents) {
    if (!parents) return 0;
    static const int32_t mano[NJ] = {-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14};
    for (int i = 0; i < NJ; ++i)
        if (parents[i] != mano[i]) return 0;
    return MB_MODEL_CHAINS_5X3;
}

// one thread per hand (mano_pose_lh.cu) once a batch fills the machine; one warp per hand below that
// [measured, pose forward + backward: B = 4096 54 us vs 82 us lane = hand; 8192 90 vs 93; 16384 136 vs 94]
static inline bool use_lane_hand(int model_flags, int mode, int B) {
    return (model_flags & MB_MODEL_CHAINS_5X3) && mode != MB_MODE_FP32 && B >= LH_MIN_HANDS;
}

extern "C" size_t mb_mano_workspace_bytes(int B, int mode) {
    if (B < 0) return 0;
    return work_layout(B, mode & 0xff).total;
}

static int check_common(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas, int B, int mode,
                        const void* workspace, size_t workspace_bytes) {
    if (B < 0 || nc < 1 || nc > NAA) return MB_E_RANGE;
    if (mode & ~(0xff | MB_MODEL_CHAINS_5X3 | MB_FWD_INFERENCE | MB_FWD_FUSED | MB_FWD_UNFUSED)) return MB_E_RANGE;
    mode &= 0xff;
    if (mode != MB_MODE_FP32 && mode != MB_MODE_F16X3 && mode != MB_MODE_F16) return MB_E_RANGE;
    if (B == 0) return 0;
    if (!blob || !rot || !coeffs || !betas) return MB_E_NULL;
    if (!workspace) return MB_E_NULL;
    if (workspace_bytes < work_layout(B, mode).total) return MB_E_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 15) || (reinterpret_cast<uintptr_t>(blob) & 15)) return MB_E_ALIGN;
    return 0;
}

// pose stage + blend contraction of the forward (fp32 FFMA or tcgen05): fills bone_t and v_posed_t
static int pose_and_blend_forward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas, int B,
                                  int mode, int model_flags, char* ws, const WorkLayout& W, float* joints, cudaStream_t s) {
    float* bone_t = reinterpret_cast<float*>(ws + W.bone_t);
    float* v_posed_t = reinterpret_cast<float*>(ws + W.v_posed_t);
    int rc;
    if (mode == MB_MODE_FP32) {
        float* feat = reinterpret_cast<float*>(ws + W.feat);
        float* rows = reinterpret_cast<float*>(ws + W.rows);
        { StageTimer t(ST_POSE_FWD, s); if ((rc = launch_pose_forward(blob, nc, rot, coeffs, betas, B, feat, nullptr, bone_t, joints, s))) return rc; }
        StageTimer t(ST_BLEND_FWD, s);
        const BlobLayout L = blob_layout();
        if ((rc = launch_sgemm(feat, FEAT_K, blob_ptr<float>(blob, L.basis), VP_PITCH, rows, VP_PITCH, B, NVC, FEAT_K, s))) return rc;
        return launch_rows_to_t(blob, rows, VP_PITCH, B, v_posed_t, s);
    }
    unsigned char* featp = reinterpret_cast<unsigned char*>(ws + W.featp);
    {
        StageTimer t(ST_POSE_FWD, s);
        rc = use_lane_hand(model_flags, mode, B) ? launch_pose_forward_lh(blob, nc, rot, coeffs, betas, B, nullptr, featp, bone_t, joints, s)
                                                 : launch_pose_forward(blob, nc, rot, coeffs, betas, B, nullptr, featp, bone_t, joints, s);
        if (rc) return rc;
    }
    StageTimer t(ST_BLEND_FWD, s);
    return launch_blend_tc_forward(blob, featp, v_posed_t, B, mode, s);
}

// the fused lane = vertex forward (vskin.cu) serves the batches of the one-thread-per-hand pose kernels — inference launches
// (5.0 against 7.7 ms per 2^20 hands) and, writing the rest-pose scratch for the backward, training launches (6.9 against 7.5);
// MB_FWD_UNFUSED selects the two separate kernels (MB_FWD_FUSED is accepted and redundant)
static inline bool use_fused_forward(int model_flags, int mode, int B) {
    return use_lane_hand(model_flags, mode, B) && !(model_flags & MB_FWD_UNFUSED);
}

// pose stage -> fused blend + skinning: verts, fingertip joints; v_posed_t only when a backward will want the workspace
static int fused_forward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas, int B, int mode,
                         int model_flags, char* ws, const WorkLayout& W, float* verts, float* joints, float* dbg, int variant,
                         cudaStream_t s) {
    float* bone_t = reinterpret_cast<float*>(ws + W.bone_t);
    unsigned char* featp = reinterpret_cast<unsigned char*>(ws + W.featp);
    float* v_posed_t = (model_flags & MB_FWD_INFERENCE) ? nullptr : reinterpret_cast<float*>(ws + W.v_posed_t);
    int rc;
    { StageTimer t(ST_POSE_FWD, s); if ((rc = launch_pose_forward_lh(blob, nc, rot, coeffs, betas, B, nullptr, featp, bone_t, joints, s))) return rc; }
    StageTimer t(ST_FUSED_FWD, s);
    // the bone operand images (1 152 B per hand) live in the backward's dv_posed tile region, which no forward kernel uses
    static_assert(VS_BONE_TILE_BYTES / VS_NH <= 74 * 16384 / 128, "bone operand images must fit the dv_posed tile region");
    unsigned char* bones_op = reinterpret_cast<unsigned char*>(ws + W.dvp);
    return launch_vskin_forward(blob, featp, bone_t, bones_op, B, mode, verts, joints, v_posed_t, dbg, variant, s);
}

extern "C" int mb_mano_forward_debug(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                                     int B, int mode, float* verts, float* joints, void* workspace, size_t workspace_bytes,
                                     float* dbg, int variant, mb_stream_t stream) {
    int rc = check_common(blob, nc, rot, coeffs, betas, B, mode, workspace, workspace_bytes);
    if (rc || B == 0) return rc;
    if (!joints || !verts) return MB_E_NULL;
    const int model_flags = mode & ~0xff;
    mode &= 0xff;
    if (!(model_flags & MB_MODEL_CHAINS_5X3) || mode == MB_MODE_FP32) return MB_E_MODEL;
    return fused_forward(blob, nc, rot, coeffs, betas, B, mode, model_flags, reinterpret_cast<char*>(workspace), work_layout(B, mode),
                         verts, joints, dbg, variant, (cudaStream_t)stream);
}

extern "C" int mb_mano_forward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                               int B, int mode, float* verts, float* joints, void* workspace, size_t workspace_bytes,
                               mb_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (verts == nullptr) {      // joints-only: no workspace needed
        if (B < 0 || nc < 1 || nc > NAA) return MB_E_RANGE;
        if (B == 0) return 0;
        if (!blob || !rot || !coeffs || !betas || !joints) return MB_E_NULL;
        StageTimer t(ST_JOINTS_FWD, s);
        if (use_lane_hand(mode & ~0xff, mode & 0xff, B)) return launch_joints_only_forward_lh(blob, nc, rot, coeffs, betas, B, joints, s);
        return launch_joints_only_forward(blob, nc, rot, coeffs, betas, B, joints, s);
    }
    int rc = check_common(blob, nc, rot, coeffs, betas, B, mode, workspace, workspace_bytes);
    if (rc || B == 0) return rc;
    if (!joints) return MB_E_NULL;
    if (reinterpret_cast<uintptr_t>(verts) & 15) return MB_E_ALIGN;
    const int model_flags = mode & ~0xff;
    mode &= 0xff;
    const WorkLayout W = work_layout(B, mode);
    char* ws = reinterpret_cast<char*>(workspace);
    if (use_fused_forward(model_flags, mode, B))
        return fused_forward(blob, nc, rot, coeffs, betas, B, mode, model_flags, ws, W, verts, joints, nullptr, 0, s);
    if ((rc = pose_and_blend_forward(blob, nc, rot, coeffs, betas, B, mode, model_flags, ws, W, joints, s))) return rc;
    StageTimer t(ST_LBS_FWD, s);
    return launch_skin_forward(blob, reinterpret_cast<float*>(ws + W.v_posed_t), reinterpret_cast<float*>(ws + W.bone_t), B,
                               verts, joints, s);
}

extern "C" int mb_mano_fit_step(const void* blob, int nc, float* params, float* exp_avg, float* exp_avg_sq,
                                const float* target_joints, const float* keypoint_vis, int B, int mode, const double* globals,
                                double* partials, float lr, float beta1, float beta2, float eps, int step, int regularize,
                                mb_stream_t stream) {
    if (B < 0 || nc < 1 || nc > NAA || step < 1) return MB_E_RANGE;
    if (!(mode & MB_MODEL_CHAINS_5X3)) return MB_E_MODEL;       // one-thread-per-hand kernel: MANO's tree only
    if (!partials) return MB_E_NULL;
    if (B == 0) return (int)cudaMemsetAsync(partials, 0, 3 * sizeof(double), (cudaStream_t)stream);
    if (!blob || !params || !exp_avg || !exp_avg_sq || !target_joints || !keypoint_vis || !globals) return MB_E_NULL;
    StageTimer t(ST_JOINTS_BWD, (cudaStream_t)stream);
    return launch_fit_step_lh(blob, nc, params, exp_avg, exp_avg_sq, target_joints, keypoint_vis, B, globals, partials, lr, beta1,
                              beta2, eps, step, regularize, (cudaStream_t)stream);
}

extern "C" int mb_mano_backward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                                const float* g_verts, const float* g_joints, int B, int mode, int flags,
                                float* g_rot, float* g_coeffs, float* g_betas, void* workspace, size_t workspace_bytes,
                                mb_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (g_verts == nullptr) {    // the heads' case: only the 21 joints carry gradient
        if (B < 0 || nc < 1 || nc > NAA) return MB_E_RANGE;
        if (B == 0) return 0;
        if (!blob || !rot || !coeffs || !betas || !g_joints || !g_rot || !g_coeffs || !g_betas) return MB_E_NULL;
        StageTimer t(ST_JOINTS_BWD, s);
        if (use_lane_hand(mode & ~0xff, mode & 0xff, B))
            return launch_joints_only_backward_lh(blob, nc, rot, coeffs, betas, g_joints, B, g_rot, g_coeffs, g_betas, s);
        return launch_joints_only_backward(blob, nc, rot, coeffs, betas, g_joints, B, g_rot, g_coeffs, g_betas, s);
    }
    int rc = check_common(blob, nc, rot, coeffs, betas, B, mode, workspace, workspace_bytes);
    if (rc || B == 0) return rc;
    if (!g_joints || !g_rot || !g_coeffs || !g_betas) return MB_E_NULL;
    const int model_flags = mode & ~0xff;
    mode &= 0xff;
    const bool lh = use_lane_hand(model_flags, mode, B);
    const WorkLayout W = work_layout(B, mode);
    char* ws = reinterpret_cast<char*>(workspace);
    float* bone_t = reinterpret_cast<float*>(ws + W.bone_t);
    float* v_posed_t = reinterpret_cast<float*>(ws + W.v_posed_t);
    float* dbone = reinterpret_cast<float*>(ws + W.dbone);
    float* dparts = reinterpret_cast<float*>(ws + W.dparts);
    float* dfeat = reinterpret_cast<float*>(ws + W.dfeat);
    const int dfeat_parts = (mode == MB_MODE_FP32) ? 1 : blend_bwd_splits(B);
    const size_t dfeat_stride = (size_t)((B + 31) / 32) * 160 * 32;
    if (!(flags & MB_BWD_WORKSPACE_VALID)) {
        // recompute the forward intermediates; joints of the recompute go to scratch (dfeat is free until step 3)
        float* scratch_joints = dfeat;      // B*63 floats <= B*148
        if ((rc = pose_and_blend_forward(blob, nc, rot, coeffs, betas, B, mode, model_flags, ws, W, scratch_joints, s))) return rc;
    }
    if (mode == MB_MODE_FP32) {
        float* dv_t = reinterpret_cast<float*>(ws + W.dv_t);
        float* rows = reinterpret_cast<float*>(ws + W.rows);
        { StageTimer t(ST_LBS_BWD, s); if ((rc = launch_skin_backward(blob, v_posed_t, bone_t, g_verts, g_joints, B, dv_t, nullptr, dbone, 0, dparts, s))) return rc; }
        StageTimer t(ST_BLEND_BWD, s);
        const BlobLayout L = blob_layout();
        if ((rc = launch_t_to_rows(blob, dv_t, VP_PITCH, B, rows, s))) return rc;
        if ((rc = launch_sgemm(rows, VP_PITCH, blob_ptr<float>(blob, L.basis_t), FEAT_K, dfeat, FEAT_K, B, FEAT_K, NVC, s))) return rc;
    } else {
        unsigned char* dvp = reinterpret_cast<unsigned char*>(ws + W.dvp);
        { StageTimer t(ST_LBS_BWD, s); if ((rc = launch_skin_backward(blob, v_posed_t, bone_t, g_verts, g_joints, B, nullptr, dvp, dbone, lh ? 1 : 0, dparts, s))) return rc; }
        StageTimer t(ST_BLEND_BWD, s);        // bf16 hi/mid x3 on tcgen05 in both tensor-core modes
        if ((rc = launch_blend_tc_backward(blob, dvp, dfeat, B, lh ? 1 : 0, dfeat_parts, dfeat_stride, s))) return rc;
    }
    StageTimer t(ST_POSE_BWD, s);
    if (lh) return launch_pose_backward_lh(blob, nc, rot, coeffs, betas, dfeat, dbone, g_joints, B, g_rot, g_coeffs, g_betas, s);
    return launch_pose_backward(blob, nc, rot, coeffs, betas, dfeat, dfeat_parts, dfeat_stride, dbone, g_joints, B, g_rot, g_coeffs, g_betas, s);
}

extern "C" size_t mb_lbs_workspace_bytes(int B) {
    if (B < 0) return 0;
    const size_t G = (size_t)((B + 31) / 32);
    return align256(sizeof(float) * G * NJ * BONE_F * 32) + align256(sizeof(float) * G * SK_NCOORD * 32);
}

extern "C" int mb_lbs_forward(const void* blob, const float* v_posed, int pitch, const float* bone, int B,
                              float* verts, float* joints, void* workspace, size_t workspace_bytes, mb_stream_t stream) {
    if (B < 0 || pitch < NVC || (pitch & 3)) return MB_E_RANGE;
    if (B == 0) return 0;
    if (!blob || !v_posed || !bone || !verts || !workspace) return MB_E_NULL;
    if (workspace_bytes < mb_lbs_workspace_bytes(B)) return MB_E_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(verts) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 15)) return MB_E_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t G = (size_t)((B + 31) / 32);
    float* bone_t = reinterpret_cast<float*>(workspace);
    float* v_posed_t = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + align256(sizeof(float) * G * NJ * BONE_F * 32));
    int rc;
    if ((rc = launch_bone_rows_to_t(bone, B, bone_t, s))) return rc;
    if ((rc = launch_rows_to_t(blob, v_posed, pitch, B, v_posed_t, s))) return rc;
    StageTimer t(ST_LBS_FWD, s);
    return launch_skin_forward(blob, v_posed_t, bone_t, B, verts, joints, s);
}
