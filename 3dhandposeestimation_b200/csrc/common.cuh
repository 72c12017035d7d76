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

// ---- "skin program" of the register-blocked skinning kernels (skin.cu) -------------------------
// Vertices are processed in blocks of 8 (24 coordinates held in registers by the lane that owns the
// hand).  Blocks are formed INSIDE 16-vertex segments of the original order (so a warp's results
// leave as contiguous 192-byte row pieces) by a host-side greedy that groups vertices with the same
// bone set; per block the program lists the distinct bones and a dense 8-vector of weights each.
// The rest-pose scratch v_posed_t is stored in this block order, hand-minor: [group][SK_NCOORD][32].
constexpr int SK_BV = 8;                                   // vertices per block
constexpr int SK_BC = SK_BV * 3;                           // coordinates per block
constexpr int SK_SEG = 16;                                 // vertices per output segment
constexpr int SK_SEG_BLKS = SK_SEG / SK_BV;                // 2
constexpr int SK_NSEG = (NV + SK_SEG - 1) / SK_SEG;        // 49
constexpr int SK_NBLK = (NV + SK_BV - 1) / SK_BV;          // 98 (the last segment has 2 blocks)
constexpr int SK_NPOS = SK_NBLK * SK_BV;                   // 784 vertex positions (6 padding)
constexpr int SK_NCOORD = SK_NPOS * 3;                     // 2352 coordinates per hand in block order
constexpr int SK_MAX_ENT = 512;                            // (block, bone) pairs supported
constexpr int SK_SLOTS = 6;                                // bone transforms a warp keeps resident in shared memory
constexpr int SK_MAX_CMD = 256;                            // slot (re)load commands per sweep
constexpr int SK_TMPL_PAD = 2400;                          // v_template in block order, padded to the GEMM's 15 x 160 columns

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
    // skin program (skin.cu)
    size_t sk_blk_ptr;  // int32 [SK_NBLK + 1]       block -> range of (block, bone) entries
    size_t sk_ent_bone; // int32 [SK_MAX_ENT]        bone | slot << 4 | wait << 7 (slot schedule, see skin.cu)
    size_t sk_cmd;      // int32 [SK_MAX_CMD + 1]    [0] = count; then (after_entry + 1) | slot << 10 | bone << 13 | next_group << 17
    size_t sk_split;    // int32 [SK_NSEG][8]        resident bones (6 slots, -1 = empty), command cursor, bones-touched mask at the start of each segment
    size_t sk_ent_w;    // float [SK_MAX_ENT][8]     dense weights of the block's 8 vertices for that bone
    size_t sk_vloc;     // uint8 [SK_NPOS]           position -> vertex index inside its 16-segment (255 = padding)
    size_t sk_perm;     // int32 [SK_NPOS]           position -> original vertex (-1 = padding)
    size_t sk_tmpl;     // float [SK_TMPL_PAD]       v_template in block (position) order
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
    L.sk_blk_ptr = o;  o = align256(o + sizeof(int32_t) * (SK_NBLK + 1));
    L.sk_ent_bone = o; o = align256(o + sizeof(int32_t) * SK_MAX_ENT);
    L.sk_cmd = o;      o = align256(o + sizeof(int32_t) * (SK_MAX_CMD + 1));
    L.sk_split = o;    o = align256(o + sizeof(int32_t) * SK_NSEG * 8);
    L.sk_ent_w = o;    o = align256(o + sizeof(float) * SK_MAX_ENT * SK_BV);
    L.sk_vloc = o;     o = align256(o + SK_NPOS);
    L.sk_perm = o;     o = align256(o + sizeof(int32_t) * SK_NPOS);
    L.sk_tmpl = o;     o = align256(o + sizeof(float) * SK_TMPL_PAD);
    L.total = o;
    return L;
}

// Workspace layout for B hands (byte offsets, 256-byte aligned sections).  "hand-minor" arrays hold
// groups of 32 hands with the hand index fastest: x_t[group][element][32].
struct WorkLayout {
    size_t bone_t;    // float [G][16][32][12]     bone transforms grouped by 32 hands: [group][bone][hand % 32][3x4]
    size_t v_posed_t; // float [G][SK_NCOORD][32]  rest-pose vertices, block order, hand-minor
    size_t dbone;     // float [B][16][12] rows or [G][16*12][32] hand-minor   (backward only)
    size_t dfeat;     // float [B][FEAT_K] rows or [G][160][32] hand-minor, x blend_bwd_splits(B) copies G*160*32 floats apart
    size_t dparts;    // float [G][units per group][16*12][32]   bone sums of split backward sweeps (small batches only)
    // tensor-core modes
    size_t featp;     // fp16 hi/lo feature tiles: ceil(B/128) * 80 KB
    size_t dvp;       // bf16 hi/mid dv_posed tiles of the tcgen05 backward: ceil(B/128) * 74 * 16 KB
    // fp32 anchor mode
    size_t feat;      // float [B][FEAT_K]
    size_t rows;      // float [B][VP_PITCH]       sgemm output (v_posed) / input (dv_posed), original vertex order
    size_t dv_t;      // float [G][SK_NCOORD][32]
    size_t total;
};

// Fewer hand groups than resident sweepers (forward: 148 x 8 warps, backward: 148 x 4 warp pairs): every
// 49-segment sweep is cut into units of 7 segments, or of 1 when even that leaves most sweepers idle.
constexpr int SKF_SWEEPERS = NUM_SMS * 8, SKB_SWEEPERS = NUM_SMS * 4;
__host__ __device__ inline int skin_segments_per_unit(long long ngroups, int sweepers) {
    // static unit -> sweeper assignment: rounds x (segments per unit + ~3 segment-times of per-unit start-up
    // [measured: B = 16384 backward, 7 units of 7 segments per pair took 224 us against 138 us unsplit]);
    // a group has ceil(49 / spu) units, the last one short
    int best = SK_NSEG;
    long long best_cost = ((ngroups + sweepers - 1) / sweepers) * SK_NSEG;
    for (int spu = SK_NSEG - 1; spu >= 1; --spu) {
        const int parts = (SK_NSEG + spu - 1) / spu;
        const long long rounds = (ngroups * parts + sweepers - 1) / sweepers;
        const long long cost = rounds * (spu + 3);
        if (cost < best_cost) { best_cost = cost; best = spu; }
    }
    return best;
}
__host__ __device__ inline int skin_units_per_group(int spu) { return (SK_NSEG + spu - 1) / spu; }
// Backward contraction at small batches: fewer CTA passes (2 x 128 hands) than SMs, so the K = 2368 loop is cut
// into up to 8 ranges; each range writes its own dfeat copy and the pose backward adds them in order.
// Only below LH_MIN_HANDS: the lane = hand pose backward reads one dfeat copy.
constexpr int LH_MIN_HANDS = 8192;   // one thread per hand pose kernels from here on, one warp per hand below
__host__ __device__ inline int blend_bwd_splits(long long B) {
    const long long passes = ((B + 127) / 128 + 1) / 2;
    if (passes <= 0 || B >= LH_MIN_HANDS) return 1;
    const long long s = NUM_SMS / passes;
    return s < 1 ? 1 : (s > 8 ? 8 : (int)s);
}
__host__ __device__ inline WorkLayout work_layout(long long B, int mode) {
    WorkLayout W;
    const size_t G = (size_t)((B + 31) / 32);
    const size_t T = (size_t)((B + 127) / 128);
    size_t o = 0;
    W.bone_t = o;    o = align256(o + sizeof(float) * G * NJ * BONE_F * 32);
    W.v_posed_t = o; o = align256(o + sizeof(float) * G * SK_NCOORD * 32);
    W.dbone = o;     o = align256(o + sizeof(float) * G * NJ * BONE_F * 32);
    W.dfeat = o;     o = align256(o + sizeof(float) * G * 160 * 32 * blend_bwd_splits(B));   // one copy per K range
    W.dparts = o;                                              // split backward sweeps: per-unit bone sums
    const int spu = skin_segments_per_unit((long long)G, SKB_SWEEPERS);
    if (spu < SK_NSEG) o = align256(o + sizeof(float) * G * skin_units_per_group(spu) * NJ * BONE_F * 32);
    W.featp = W.dvp = W.feat = W.rows = W.dv_t = o;
    if (mode == MB_MODE_FP32) {
        W.feat = o;  o = align256(o + sizeof(float) * B * FEAT_K);
        W.rows = o;  o = align256(o + sizeof(float) * B * VP_PITCH);
        W.dv_t = o;  o = align256(o + sizeof(float) * G * SK_NCOORD * 32);
    } else {
        W.featp = o; o = align256(o + T * 81920);
        W.dvp = o;   o = align256(o + T * (74 * 16384));
    }
    W.total = o;
    return W;
}

template <typename T>
__host__ __device__ inline const T* blob_ptr(const void* blob, size_t off) {
    return reinterpret_cast<const T*>(reinterpret_cast<const char*>(blob) + off);
}

// ---- kernel launchers implemented in the other translation units ----------
// feat (fp32 rows) and featp (fp16 hi/lo UMMA tiles) may each be NULL
// bone_t[group][bone][hand % 32][12]
int launch_pose_forward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                        int B, float* feat, unsigned char* featp, float* bone_t, float* joints, cudaStream_t s);
// one-thread-per-hand variants (mano_pose_lh.cu): MANO tree only; dfeat_t [G][160][32], dbone_t [G][192][32] hand-minor
int launch_pose_forward_lh(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                           int B, float* feat, unsigned char* featp, float* bone_t, float* joints, cudaStream_t s);
int launch_pose_backward_lh(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                            const float* dfeat_t, const float* dbone_t, const float* g_joints, int B,
                            float* g_rot, float* g_coeffs, float* g_betas, cudaStream_t s);
int launch_joints_only_forward_lh(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                                  int B, float* joints, cudaStream_t s);
int launch_joints_only_backward_lh(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                                   const float* g_joints, int B, float* g_rot, float* g_coeffs, float* g_betas, cudaStream_t s);
int launch_fit_step_lh(const void* blob, int nc, float* params, float* exp_avg, float* exp_avg_sq, const float* target_joints,
                       const float* vis, int B, const double* globals, double* partials, float lr, float beta1, float beta2,
                       float eps, int step, int regularize, cudaStream_t s);
// dfeat: dfeat_parts partial copies, dfeat_stride floats apart, added in order
int launch_pose_backward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                         const float* dfeat, int dfeat_parts, size_t dfeat_stride, const float* dbone, const float* g_joints,
                         int B, float* g_rot, float* g_coeffs, float* g_betas, cudaStream_t s);
int launch_joints_only_forward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                               int B, float* joints, cudaStream_t s);
int launch_joints_only_backward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                                const float* g_joints, int B, float* g_rot, float* g_coeffs, float* g_betas,
                                cudaStream_t s);
// C[M][ldc] = A[M][lda] (K cols) * Bm[K][ldb] (N cols), fp32 FFMA
int launch_sgemm(const float* A, int lda, const float* Bm, int ldb, float* C, int ldc,
                 long long M, int N, int K, cudaStream_t s);
// launch bookkeeping (api.cu): every kernel launch of this library goes through cuda_rc()
void count_launch();
inline int cuda_rc() {
    count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: a process that drives several GPUs must
// set it on each of them (a process-wide "done" flag made the first launch on a second device fail with
// invalid-value).  One bit per device ordinal, lock-free; setting it twice is harmless.
struct SmemAttrOnce { unsigned long long mask[4]; };
template <class Kernel>
inline int ensure_dyn_smem(SmemAttrOnce& once, Kernel kernel, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    unsigned long long* word = &once.mask[(dev >> 6) & 3];
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(word, __ATOMIC_ACQUIRE) & bit) return 0;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    __atomic_fetch_or(word, bit, __ATOMIC_RELEASE);
    return 0;
}

// per-stage CUDA-event profiling (off by default; bench.py turns it on for the roofline pass)
enum Stage { ST_POSE_FWD = 0, ST_BLEND_FWD, ST_LBS_FWD, ST_LBS_BWD, ST_BLEND_BWD, ST_POSE_BWD,
             ST_JOINTS_FWD, ST_JOINTS_BWD, ST_FK_FWD, ST_FK_BWD, ST_FUSED_FWD, ST_COUNT };
struct StageTimer {
    StageTimer(int stage, cudaStream_t s);
    ~StageTimer();
    int stage; cudaStream_t stream; void* rec;
};

}  // namespace mb
