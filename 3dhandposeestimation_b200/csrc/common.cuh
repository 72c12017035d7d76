// common.cuh — shared definitions of libmano_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mano_b200.h"

#define HD __host__ __device__ __forceinline__

namespace mb {

constexpr int NV = 778;           // vertices
constexpr int NVC = 2334;         // vertex coordinates
constexpr int VP_PITCH = 2336;    // row pitch (floats) of v_posed / dv_posed scratch: 16-byte aligned rows
constexpr int NJ = 16;            // chain joints
constexpr int NOUTJ = 21;         // output joints
constexpr int NB = 10;            // betas
constexpr int NPF = 135;          // pose-feature length
constexpr int NAA = 45;           // articulated axis-angle length
constexpr int FEAT_K = 148;       // [beta | pf | 1 | 0 0]
constexpr int FEAT_ONE = 145;
constexpr int MAX_INFL = 8;
constexpr int BONE_F = 12;        // 3x4 [R|t] per bone
constexpr int NUM_SMS = 148;
constexpr int MAX_NNZ = 3200;      // total skinning weights supported
constexpr int LBS_WARPS = 16;      // warps per CTA of the skinning kernels
constexpr int LBS_CV = 64;         // vertices per chunk
constexpr int LBS_CF = LBS_CV * 3; // floats per chunk row
constexpr int LBS_CHUNKS = (NV + LBS_CV - 1) / LBS_CV;   // 13
constexpr int LBS_SLOTS = 2;       // bones (or parts of long bones) owned by one warp in the backward reduction

// Device blob layout (byte offsets, every section 256-byte aligned).
struct BlobLayout {
    size_t header;      // BlobHeader
    size_t basis;       // float [FEAT_K][VP_PITCH]   (row pitch padded to 2336)
    size_t basis_t;     // float [NVC][FEAT_K]
    size_t j0;          // float [16][3]
    size_t jb;          // float [16][3][10]
    size_t pca;         // float [45][45] (first nc rows valid)
    size_t pose_mean;   // float [45]
    size_t skin_w;      // float [778][8]
    size_t skin_b;      // uint8 [778][8]
    size_t skin_cnt;    // uint8 [778] (+pad)
    size_t csc_ptr;     // int32 [17]      bone -> range in csc_v / csc_w
    size_t csc_v;       // int32 [778*8]   vertex ids grouped by bone
    size_t csc_w;       // float [778*8]
    // tables of the lane=hand skinning kernels (mano_lbs.cu)
    size_t csr_ptr;     // int32 [779]     vertex -> range in csr_w / csr_b
    size_t csr_w;       // float [MAX_NNZ]
    size_t csr_b;       // uint8 [MAX_NNZ]
    size_t bseg;        // int32 [LBS_WARPS][LBS_CHUNKS][LBS_SLOTS][2]  entry ranges per (warp, vertex chunk, owned slot)
    size_t bent_idx;    // uint16 [MAX_NNZ] float index of the entry's vertex inside its chunk
    size_t bent_w;      // float  [MAX_NNZ]
    size_t bslot;       // int32 [LBS_WARPS][LBS_SLOTS] owned bone id (-1 = none)
    size_t total;
};

struct BlobHeader {
    int32_t magic;          // 'MB20'
    int32_t abi;
    int32_t nc;
    int32_t max_depth;
    int32_t parents[NJ];
    int32_t depth[NJ];
    int32_t n_children[NJ];
    int32_t children[NJ][NJ];
    int32_t max_children_at_depth[NJ];   // max #children of any node at depth d
    int32_t csc_nnz;
};

__host__ __device__ constexpr size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

__host__ __device__ inline BlobLayout blob_layout() {
    BlobLayout L;
    size_t o = 0;
    L.header = o;    o = align256(o + sizeof(BlobHeader));
    L.basis = o;     o = align256(o + sizeof(float) * FEAT_K * VP_PITCH);
    L.basis_t = o;   o = align256(o + sizeof(float) * NVC * FEAT_K);
    L.j0 = o;        o = align256(o + sizeof(float) * NJ * 3);
    L.jb = o;        o = align256(o + sizeof(float) * NJ * 3 * NB);
    L.pca = o;       o = align256(o + sizeof(float) * NAA * NAA);
    L.pose_mean = o; o = align256(o + sizeof(float) * NAA);
    L.skin_w = o;    o = align256(o + sizeof(float) * NV * MAX_INFL);
    L.skin_b = o;    o = align256(o + NV * MAX_INFL);
    L.skin_cnt = o;  o = align256(o + NV);
    L.csc_ptr = o;   o = align256(o + sizeof(int32_t) * (NJ + 1));
    L.csc_v = o;     o = align256(o + sizeof(int32_t) * NV * MAX_INFL);
    L.csc_w = o;     o = align256(o + sizeof(float) * NV * MAX_INFL);
    L.csr_ptr = o;   o = align256(o + sizeof(int32_t) * (NV + 1));
    L.csr_w = o;     o = align256(o + sizeof(float) * MAX_NNZ);
    L.csr_b = o;     o = align256(o + MAX_NNZ);
    L.bseg = o;      o = align256(o + sizeof(int32_t) * LBS_WARPS * LBS_CHUNKS * LBS_SLOTS * 2);
    L.bent_idx = o;  o = align256(o + sizeof(uint16_t) * MAX_NNZ);
    L.bent_w = o;    o = align256(o + sizeof(float) * MAX_NNZ);
    L.bslot = o;     o = align256(o + sizeof(int32_t) * LBS_WARPS * LBS_SLOTS);
    L.total = o;
    return L;
}

// Workspace layout for B hands (byte offsets, 256-byte aligned sections).
struct WorkLayout {
    size_t feat;      // float [B][FEAT_K]
    size_t bone;      // float [B][16][12]
    size_t v_posed;   // float [B][VP_PITCH]
    size_t dv_posed;  // ALIASES v_posed: the skinning backward overwrites each row after reading it
    size_t dbone;     // float [B][16][12]     (backward only)
    size_t dfeat;     // float [B][FEAT_K]     (backward only)
    size_t featp;     // fp16 hi/lo feature tiles of the tcgen05 path: ceil(B/128) * 80 KB
    size_t dvp;       // bf16 hi/mid dv_posed tiles of the tcgen05 backward: ceil(B/128) * 73 * 16 KB
    size_t total;
};

__host__ __device__ inline WorkLayout work_layout(long long B) {
    WorkLayout W;
    size_t o = 0;
    W.feat = o;     o = align256(o + sizeof(float) * B * FEAT_K);
    W.bone = o;     o = align256(o + sizeof(float) * B * NJ * BONE_F);
    W.v_posed = o;  o = align256(o + sizeof(float) * B * VP_PITCH);
    W.dv_posed = W.v_posed;
    W.dbone = o;    o = align256(o + sizeof(float) * B * NJ * BONE_F);
    W.dfeat = o;    o = align256(o + sizeof(float) * B * FEAT_K);
    W.featp = o;    o = align256(o + (size_t)((B + 127) / 128) * 81920);
    W.dvp = o;      o = align256(o + (size_t)((B + 127) / 128) * (73 * 16384));
    W.total = o;
    return W;
}

template <typename T>
__host__ __device__ inline const T* blob_ptr(const void* blob, size_t off) {
    return reinterpret_cast<const T*>(reinterpret_cast<const char*>(blob) + off);
}

// ---- kernel launchers implemented in the other translation units ----------
// feat (fp32 rows) and featp (fp16 hi/lo UMMA tiles) may each be NULL
int launch_pose_forward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                        int B, float* feat, unsigned char* featp, float* bone, float* joints, cudaStream_t s);
int launch_pose_backward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                         const float* dfeat, const float* dbone, const float* g_joints, int B,
                         float* g_rot, float* g_coeffs, float* g_betas, cudaStream_t s);
int launch_joints_only_forward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                               int B, float* joints, cudaStream_t s);
int launch_joints_only_backward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                                const float* g_joints, int B, float* g_rot, float* g_coeffs, float* g_betas,
                                cudaStream_t s);
// C[M][ldc] = A[M][lda] (K cols) * Bm[K][ldb] (N cols), fp32 FFMA
int launch_sgemm(const float* A, int lda, const float* Bm, int ldb, float* C, int ldc,
                 long long M, int N, int K, cudaStream_t s);
int launch_lbs_forward(const void* blob, const float* v_posed, int pitch, const float* bone, int B,
                       float* verts, float* joints, cudaStream_t s);
// dv_posed (fp32 rows) or dvp (bf16 hi/mid UMMA tiles for the tcgen05 backward): exactly one is non-NULL
int launch_lbs_backward(const void* blob, const float* v_posed, int pitch, const float* bone,
                        const float* g_verts, const float* g_joints, int B,
                        float* dv_posed, unsigned char* dvp, float* dbone, cudaStream_t s);

// launch bookkeeping (api.cu): every kernel launch of this library goes through cuda_rc()
void count_launch();
inline int cuda_rc() {
    count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

// per-stage CUDA-event profiling (off by default; bench.py turns it on for the roofline pass)
enum Stage { ST_POSE_FWD = 0, ST_BLEND_FWD, ST_LBS_FWD, ST_LBS_BWD, ST_BLEND_BWD, ST_POSE_BWD,
             ST_JOINTS_FWD, ST_JOINTS_BWD, ST_FK_FWD, ST_FK_BWD, ST_COUNT };
struct StageTimer {
    StageTimer(int stage, cudaStream_t s);
    ~StageTimer();
    int stage; cudaStream_t stream; void* rec;
};

}  // namespace mb
