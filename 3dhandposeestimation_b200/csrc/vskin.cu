// vskin.cu — blend shapes + linear blend skinning of the forward pass in ONE kernel, LANE = VERTEX, both
// contractions on the 5th-generation tensor cores (tcgen05), no rest-pose scratch in HBM.
//
// Reference: MANOLayer.py:130-137 (v_posed = v_template + S beta + P pf), :177-185 (T_v = sum_k w_vk A_k,
// v' = T_v [v_posed; 1]), :190-202 (fingertip vertices -> joints 4, 8, 12, 16, 20), :188/:204-205 (global rotation,
// folded into the bone transforms by the pose stage).
//
// Why another mapping (profiles/r1 -> r2): with lane = hand (skin.cu) the weight blend sum_k w_vk A_k costs
// ~31 FMAs per (vertex, hand) on the CUDA cores plus a transposition of every result through shared memory —
// 1 750 warp instructions per hand, issue-bound at ~5 ms per 2^20 hands — and the rest-pose vertices cross HBM
// twice (written by the blend GEMM, read here).  Here both contractions are MMAs with M = 128 VERTICES:
//   v_posed[v][h] = sum_f basis[v][f] feat[h][f]      three planes x, y, z;  N = 64 hands, K = 160, fp16 hi/lo x 3 products
//   T[v][(h, e)]  = sum_k W[v][k] A[h][k][e]          N = 4 hands x 12 elements, K = 16 bones, fp16 3-way split, 4 products
// Both land in TMEM with lane = vertex; the epilogue thread of vertex v reads its rest position (3 columns) and its
// blended 3x4 transform (12 columns) of a hand, does 12 FMAs, and the warp's 32 consecutive vertices leave through a
// shared-memory row as ONE contiguous, sector-aligned 1 536-byte piece per (hand, vertex tile) of
// verts[B][778][3] in its natural layout — no transposition, no v_posed_t round trip.
//
// Roles (512 threads, 1 CTA per SM, persistent over 64-hand tiles):
//   warp 0   basis producer: (tile, plane, K chunk) stages of 16 KB, always L2 hits, 4-stage ring
//   warp 1   MMA issuer (one thread): blend products of vertex tile t+1 interleaved with the transform chunks of tile t
//   warp 2   TMEM allocation; producer of the hand tile's feature rows (40 KB) and of the weight tiles (12 KB per
//            vertex tile, double buffered)
//   warp 3   converts the hand tile's fp32 bone transforms (48 KB, straight from the pose stage's bone_t) into the
//            transform products' fp16 x3 B operand in shared memory (72 KB)
//   warps 12-15  store warps, one per slot of the result-row ring
//   warps 4-11  epilogue: warp % 4 = TMEM lane quarter (32 vertices); warps 4-7 take the even chunks, 8-11 the odd ones
// TMEM (512 columns): two rest-position stages of 3 x 64 columns, two transform stages of 48 columns.
#include <cuda_fp16.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include "common.cuh"
#include "blend_tc.cuh"
#include "vskin.cuh"
#include "ptx.cuh"
#include "tc_ptx.cuh"

namespace mb {
namespace {

constexpr int VS_THREADS = 512;                   // 16 warps: see the role list above
constexpr int VS_EPI_WARPS = 8;
constexpr int VS_ASTAGES = 4;
constexpr int VS_OSTAGES = 4;                     // result-row ring: two chunks per epilogue warp set
constexpr int VS_ROW = 392;                        // floats per staging row: up to 6 carried floats + 384 + slack
constexpr uint32_t VS_TMEM_COLS = 512;
constexpr uint32_t VS_VP_COLS = 3 * VS_NH;         // 192 columns per rest-position stage (two stages)
constexpr uint32_t VS_T_COL0 = 2 * VS_VP_COLS;     // 384: first transform column
constexpr int VS_TSTAGES = 2;                      // transform stages of 48 columns, one per epilogue warp set: 480 columns in all
// instruction descriptors: f16 x f16 -> f32, M = 128
constexpr uint32_t VS_IDESC_BLEND = (1u << 4) | ((uint32_t)(VS_NH >> 3) << 17) | ((uint32_t)(VS_M >> 4) << 24);                // A, B K-major
constexpr uint32_t VS_IDESC_T = (1u << 4) | (1u << 16) | ((uint32_t)(VS_TN >> 3) << 17) | ((uint32_t)(VS_M >> 4) << 24);       // B MN-major

__constant__ int c_vs_tip_vert[5] = {333, 444, 672, 555, 745};
__constant__ int c_vs_tip_slot[5] = {4, 8, 12, 16, 20};

struct VsShared {
    alignas(128) unsigned char feat[TC_K_CHUNKS][2][VS_NH * TC_K_CHUNK * 2];      // 40 KB: [K chunk][hi, lo][64 hands x 32]
    alignas(128) unsigned char bones[VS_NCH][VS_BONE_SPLITS][VS_BONE_CHUNK_BYTES]; // 72 KB
    alignas(128) unsigned char a[VS_ASTAGES][VS_A_STAGE_BYTES];                    // 64 KB basis ring
    alignas(128) unsigned char w[2][VS_W_TILE_BYTES];                              // 24 KB weight tiles
    alignas(128) float out[VS_OSTAGES][VS_HC][VS_ROW];                             // 18.4 KB result rows
    alignas(16) float carry[2][VS_NH][8];                                          // tail of a hand's row piece, for the next vertex tile
    alignas(8) unsigned long long a_full[VS_ASTAGES], a_empty[VS_ASTAGES];
    unsigned long long w_full[2], w_empty[2];
    unsigned long long feat_full, feat_empty, bones_full, bones_empty;
    unsigned long long vp_full[2], vp_empty[2];
    unsigned long long t_full[VS_TSTAGES], t_empty[VS_TSTAGES];
    unsigned long long out_full[VS_OSTAGES], out_empty[VS_OSTAGES];
    uint32_t tmem_base;
};

// A wait that cannot hang the GPU: a broken pipeline traps after ~1 s instead of spinning forever.
__device__ __forceinline__ void vs_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t b = smem_u32(bar);
    if (mbar_try_wait(b, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(b, parity)) {
        if (clock64() - t0 > 2000000000LL) __trap();
    }
}

// sector alignment of the row pieces: rows of verts[B][778][3] are 9 336 B = 24 (mod 32) apart, so the piece of
// vertex tile t >= 1 of hand h covers floats [384 t - d, 384 t + 384 - d), d = (0, 6, 4, 2)[h % 4]: every piece starts on
// a 32-byte boundary; the d floats in front are the tail of the previous tile's results (carry).  Tile 0 starts at the
// row's first 32-byte boundary (float a = (8 - d) % 8) and its first a floats are plain stores; inside the shared row
// the results sit S floats in so that the bulk copy's source is 16-byte aligned: S = d for t >= 1, (0, 2, 0, 2)[h % 4] for t = 0.
__device__ __forceinline__ int vs_d(int hl) { return (8 - 2 * hl) & 7; }                 // hl = h % 4 -> 0, 6, 4, 2
__device__ __forceinline__ int vs_shift(int t, int hl) { return t == 0 ? ((hl & 1) << 1) : vs_d(hl); }

__global__ void __launch_bounds__(VS_THREADS, 1)
vskin_forward_kernel(const TcBlobHeader* __restrict__ hdr, const unsigned char* __restrict__ vs_basis,
                     const unsigned char* __restrict__ vs_w, const float4* __restrict__ vs_tmpl,
                     const unsigned char* __restrict__ featp, const float* __restrict__ bone_t,
                     int B, int ntiles, int blend_products, int t_products,
                     float* __restrict__ verts, float* __restrict__ joints, float* __restrict__ v_posed_t,
                     float* __restrict__ dbg, int variant) {
    extern __shared__ unsigned char smem_raw[];
    VsShared& S = *reinterpret_cast<VsShared*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Thread-block cluster (1, 2 or 4 CTAs): the CTAs of a cluster walk DIFFERENT hand tiles in lockstep and share one
    // basis stream — each loads 1 / csize of every stage and multicasts it to all of them.  [profiles/r2: unicast, the
    // basis stream alone (1.68 MB per 64-hand tile = 26 KB per hand, every SM reading the same L2 lines) ran the kernel
    // at ~15 B/clk/SM: 6.7 ms per 2^20 hands with every MMA, store and FMA removed]
    const uint32_t csize = cluster_nctarank(), crank = cluster_ctarank();
    const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
    const int nclusters = (int)(gridDim.x / csize), cluster_id = (int)(blockIdx.x / csize);
    const int nrounds = (ntiles + (int)csize - 1) / (int)csize;         // a round = one hand tile per CTA of the cluster
    // a CTA whose tile of the last round does not exist runs the round on tile `ntiles - 1`'s operands and stores nothing
#define VS_FOR_EACH_TILE for (int rnd = cluster_id, tile = rnd * (int)csize + (int)crank; rnd < nrounds; rnd += nclusters, tile = rnd * (int)csize + (int)crank)

    if (threadIdx.x == 0) {
        for (int s = 0; s < VS_ASTAGES; ++s) { mbar_init(smem_u32(&S.a_full[s]), 1); mbar_init(smem_u32(&S.a_empty[s]), csize); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&S.w_full[s]), 1); mbar_init(smem_u32(&S.w_empty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&S.vp_full[s]), 1); mbar_init(smem_u32(&S.vp_empty[s]), VS_EPI_WARPS); }
        for (int s = 0; s < VS_TSTAGES; ++s) { mbar_init(smem_u32(&S.t_full[s]), 1); mbar_init(smem_u32(&S.t_empty[s]), VS_EPI_WARPS / 2); }
        mbar_init(smem_u32(&S.feat_full), 1); mbar_init(smem_u32(&S.feat_empty), 1);
        mbar_init(smem_u32(&S.bones_full), 1); mbar_init(smem_u32(&S.bones_empty), 1);
        for (int s = 0; s < VS_OSTAGES; ++s) { mbar_init(smem_u32(&S.out_full[s]), VS_EPI_WARPS / 2); mbar_init(smem_u32(&S.out_empty[s]), 1); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(&S.tmem_base), VS_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                              // every CTA's barriers exist before a peer signals them
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;

    if (warp == 0) {
        if (elect_one()) {
            // ===== basis producer: the same 105-stage stream for every hand tile, kept in L2 =====
            const uint64_t keep = l2_policy_evict_last();
            uint32_t stage = 0, phase = 0;
            const uint32_t slice = VS_A_STAGE_BYTES / csize;           // this CTA's share of every stage
            VS_FOR_EACH_TILE {
                (void)tile;
                if (variant & 0x8000) continue;                        // experiment: no basis stream
                for (int i = 0; i < VS_NT * VS_STAGES_PER_TILE; ++i) {
                    vs_wait(&S.a_empty[stage], phase ^ 1);             // every CTA of the cluster has consumed the slot
                    mbar_expect_tx(smem_u32(&S.a_full[stage]), VS_A_STAGE_BYTES);
                    if (csize == 1)
                        bulk_g2s_hint(smem_u32(S.a[stage]), vs_basis + (size_t)i * VS_A_STAGE_BYTES, VS_A_STAGE_BYTES,
                                      smem_u32(&S.a_full[stage]), keep);
                    else
                        bulk_g2s_multicast(smem_u32(S.a[stage]) + crank * slice, vs_basis + (size_t)i * VS_A_STAGE_BYTES + crank * slice, slice,
                                           smem_u32(&S.a_full[stage]), cmask, keep);
                    if (++stage == VS_ASTAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        if (elect_one()) {
            // ===== per-hand-tile operands (features, bones) and the weight tiles =====
            const uint64_t keep = l2_policy_evict_last(), once = l2_policy_evict_first();
            uint32_t it = 0, gw = 0;
            VS_FOR_EACH_TILE {
                const int ltile = tile < ntiles ? tile : ntiles - 1;
                vs_wait(&S.feat_empty, (it & 1) ^ 1);
                mbar_expect_tx(smem_u32(&S.feat_full), TC_K_CHUNKS * 2 * 4096);
                const unsigned char* fsrc = featp + (size_t)(ltile >> 1) * TC_A_TILE_BYTES + (size_t)(ltile & 1) * 4096;
                for (int c = 0; c < TC_K_CHUNKS; ++c)
                    for (int sp = 0; sp < 2; ++sp)
                        bulk_g2s_hint(smem_u32(S.feat[c][sp]), fsrc + (size_t)c * TC_A_STAGE_BYTES + (size_t)sp * TC_A_BLOCK_BYTES, 4096,
                                      smem_u32(&S.feat_full), once);
                for (int t = 0; t < VS_NT; ++t, ++gw) {
                    vs_wait(&S.w_empty[gw & 1], ((gw >> 1) & 1) ^ 1);
                    mbar_expect_tx(smem_u32(&S.w_full[gw & 1]), VS_W_TILE_BYTES);
                    bulk_g2s_hint(smem_u32(S.w[gw & 1]), vs_w + (size_t)t * VS_W_TILE_BYTES, VS_W_TILE_BYTES, smem_u32(&S.w_full[gw & 1]), keep);
                }
                ++it;
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ===== MMA issuer =====
            // [profiles/r2: issued by `lane == 0` the compiler wrapped every tcgen05.mma in a divergence "waterfall" (ELECT /
            // R2UR.BROADCAST / BRA.U.ANY) and rebuilt 64-bit descriptors per instruction — ~350 scalar instructions per
            // 4-hand chunk at IPC ~0.25: the issuing THREAD, not the tensor pipe (14 % busy), paced the kernel at ~1 500 clk
            // per chunk.  Now: elect.sync, 32-bit descriptor halves (only the low word changes), the chunk loop unrolled so
            // that plane / K-chunk / accumulate flags are constants.]
            uint32_t a_stage = 0, a_phase = 0, gt = 0, it = 0;
            // K-major operands: LBO 128 B (between the K core matrices of one MMA), SBO between 8-row groups
            const uint32_t hi_k512 = desc_hi(TC_SBO), hi_k256 = desc_hi(256);
            const uint32_t a_lo0 = desc_lo(smem_u32(S.a[0]), TC_LBO), f_lo0 = desc_lo(smem_u32(S.feat[0][0]), TC_LBO);
            const uint32_t w_lo0 = desc_lo(smem_u32(S.w[0]), 128);
            // bones: MN-major, n-groups 256 B apart (SBO), k-groups 128 B apart (LBO); variant 1 swaps the two fields
            const uint32_t b_lo0 = desc_lo(smem_u32(&S.bones[0][0][0]), (variant & 1) ? 256 : 128);
            const uint32_t hi_b = desc_hi((variant & 1) ? 128 : 256);
            // one (plane p, K chunk c) stage of the blend products of vertex-tile counter g
            auto blend_stage = [&](uint32_t g, int p, int c) {
                if (p == 0 && c == 0) { vs_wait(&S.vp_empty[g & 1], ((g >> 1) & 1) ^ 1); tc_fence_after(); }
                if (variant & 0x8000) {                                 // experiment: no basis stream
                    if (p == 2 && c == TC_K_CHUNKS - 1) tc_commit(smem_u32(&S.vp_full[g & 1]));
                    return;
                }
                vs_wait(&S.a_full[a_stage], a_phase);
                tc_fence_after();
                const uint32_t d = tmem + (g & 1) * VS_VP_COLS + p * VS_NH;
                const uint32_t a_st = a_lo0 + ((a_stage * VS_A_STAGE_BYTES) >> 4);
                const uint32_t f_st = f_lo0 + ((c * 2 * 4096) >> 4);
#pragma unroll
                for (int j = 0; j < TC_K_CHUNK / 16; ++j) {
                    const uint32_t a_hi = a_st + ((j * 2 * (int)TC_LBO) >> 4), a_lo = a_hi + ((VS_A_STAGE_BYTES / 2) >> 4);
                    const uint32_t b_hi = f_st + ((j * 2 * (int)TC_LBO) >> 4), b_lo = b_hi + (4096 >> 4);
                    if (variant & 0x400) continue;                                  // experiment: no blend products
                    if (c == 0 && j == 0) umma_f16_lohi<false>(d, a_hi, hi_k512, b_hi, hi_k512, VS_IDESC_BLEND);
                    else umma_f16_lohi<true>(d, a_hi, hi_k512, b_hi, hi_k512, VS_IDESC_BLEND);
                    if (blend_products == 3) {
                        umma_f16_lohi<true>(d, a_lo, hi_k512, b_hi, hi_k512, VS_IDESC_BLEND);
                        umma_f16_lohi<true>(d, a_hi, hi_k512, b_lo, hi_k512, VS_IDESC_BLEND);
                    }
                }
                if (csize == 1) tc_commit(smem_u32(&S.a_empty[a_stage]));
                else tc_commit_multicast(smem_u32(&S.a_empty[a_stage]), cmask);      // every producer of the cluster refills this slot
                if (++a_stage == VS_ASTAGES) { a_stage = 0; a_phase ^= 1; }
                if (p == 2 && c == TC_K_CHUNKS - 1) tc_commit(smem_u32(&S.vp_full[g & 1]));
            };
            VS_FOR_EACH_TILE {
                (void)tile;
                vs_wait(&S.feat_full, it & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < VS_STAGES_PER_TILE; ++k) blend_stage(gt, k / TC_K_CHUNKS, k % TC_K_CHUNKS);
                vs_wait(&S.bones_full, it & 1);
                tc_fence_after();
#pragma unroll 1
                for (int t = 0; t < VS_NT; ++t, ++gt) {
                    vs_wait(&S.w_full[gt & 1], (gt >> 1) & 1);
                    tc_fence_after();
                    const uint32_t w1 = w_lo0 + (((gt & 1) * VS_W_TILE_BYTES) >> 4), w2 = w1 + (4096 >> 4);
                    const bool more = t + 1 < VS_NT;
#pragma unroll
                    for (int ch = 0; ch < VS_NCH; ++ch) {
                        // T stage ch & 1: its uses so far = 8 gt + ch / 2 -> wait parity ((ch >> 1) & 1) ^ 1 (8 per tile: even)
                        vs_wait(&S.t_empty[ch & 1], (((ch >> 1) & 1) ^ 1));
                        tc_fence_after();
                        const uint32_t d = tmem + VS_T_COL0 + (ch & 1) * VS_TN;
                        const uint32_t a1 = b_lo0 + ((ch * VS_BONE_SPLITS * VS_BONE_CHUNK_BYTES) >> 4);
                        const uint32_t a2 = a1 + (VS_BONE_CHUNK_BYTES >> 4), a3 = a2 + (VS_BONE_CHUNK_BYTES >> 4);
                        // smallest products first, w1 a1 last: the tensor core TRUNCATES its fp32 accumulator at the running
                        // sum's magnitude on every MMA; with the corrections accumulated among themselves first, only the
                        // last addition rounds at full magnitude
                        if (!(variant & 0x800)) {                       // 0x800: experiment, no transform products
                            if (t_products >= 4) {
                                umma_f16_lohi<false>(d, w1, hi_k256, a3, hi_b, VS_IDESC_T);
                                if (t_products >= 5) umma_f16_lohi<true>(d, w2, hi_k256, a2, hi_b, VS_IDESC_T);
                                umma_f16_lohi<true>(d, w1, hi_k256, a2, hi_b, VS_IDESC_T);
                                umma_f16_lohi<true>(d, w2, hi_k256, a1, hi_b, VS_IDESC_T);
                                umma_f16_lohi<true>(d, w1, hi_k256, a1, hi_b, VS_IDESC_T);
                            } else if (t_products == 3) {
                                umma_f16_lohi<false>(d, w1, hi_k256, a2, hi_b, VS_IDESC_T);
                                umma_f16_lohi<true>(d, w2, hi_k256, a1, hi_b, VS_IDESC_T);
                                umma_f16_lohi<true>(d, w1, hi_k256, a1, hi_b, VS_IDESC_T);
                            } else {
                                umma_f16_lohi<false>(d, w1, hi_k256, a1, hi_b, VS_IDESC_T);
                            }
                        }
                        tc_commit(smem_u32(&S.t_full[ch & 1]));
                        // the next vertex tile's blend products, one stage per chunk (15 stages over 16 chunks)
                        if (more && ch < VS_STAGES_PER_TILE) blend_stage(gt + 1, ch / TC_K_CHUNKS, ch % TC_K_CHUNKS);
                    }
                    tc_commit(smem_u32(&S.w_empty[gt & 1]));
                    if (t == VS_NT - 2) tc_commit(smem_u32(&S.feat_empty));        // blend products of the last vertex tile are issued
                }
                tc_commit(smem_u32(&S.bones_empty));
                ++it;
            }
        }
    } else if (warp == 3) {
        // ===== bone operand: fp32 transforms bone_t[group][bone][hand % 32][12] -> fp16 x3 MN-major core matrices =====
        // The 12 elements of 4 consecutive hands of one bone are 48 contiguous floats = one K row (k = bone) of a chunk's
        // B operand, n = (hand % 4) * 12 + element: a lane converts 8 of them (one 16-byte n-group) into the three splits
        // a = a1 + a2 + a3 of 2^4 a.  [profiles/r2: written by the pose kernel as 96 scattered 16 / 8-byte stores per
        // hand it cost that kernel +1.6 ms per 2^20 hands and 2.3 KB per hand of HBM traffic]
        const long long ngroups = ((long long)B + 31) >> 5;
        uint32_t it = 0;
        VS_FOR_EACH_TILE {
            vs_wait(&S.bones_empty, (it & 1) ^ 1);
#pragma unroll 1
            for (int ch = 0; ch < VS_NCH; ++ch) {
                const long long group = (long long)tile * 2 + (ch >> 3);
                const float* src0 = bone_t + (size_t)group * (NJ * BONE_F * 32) + 48 * (ch & 7);
                float4 v[3][2];
#pragma unroll
                for (int r = 0; r < 3; ++r) {                          // 96 (bone, n-group) items per chunk: three per lane
                    const int i = lane + 32 * r, k = i / 6, g = i - 6 * k;
                    const float4* p = reinterpret_cast<const float4*>(src0 + (size_t)k * (BONE_F * 32) + 8 * g);
                    if (group < ngroups) { v[r][0] = __ldg(p); v[r][1] = __ldg(p + 1); }
                    else { v[r][0] = make_float4(0.f, 0.f, 0.f, 0.f); v[r][1] = v[r][0]; }
                }
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int i = lane + 32 * r, k = i / 6, g = i - 6 * k;
                    float x[8] = {v[r][0].x, v[r][0].y, v[r][0].z, v[r][0].w, v[r][1].x, v[r][1].y, v[r][1].z, v[r][1].w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) x[e] *= (float)(1 << VS_BONE_SCALE_LOG2);
                    unsigned char* dst = &S.bones[ch][0][0] + g * 256 + (k >> 3) * 128 + (k & 7) * 16;
#pragma unroll
                    for (int sp = 0; sp < VS_BONE_SPLITS; ++sp) {
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __half lo = __float2half_rn(x[2 * e]), hi = __float2half_rn(x[2 * e + 1]);
                            x[2 * e] -= __half2float(lo);
                            x[2 * e + 1] -= __half2float(hi);
                            pk[e] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
                        }
                        *reinterpret_cast<uint4*>(dst + sp * VS_BONE_CHUNK_BYTES) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
            }
            fence_proxy_async();                                       // generic-proxy writes -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&S.bones_full));
            ++it;
        }
    } else if (warp >= 12) {
        // ===== store warps: result rows shared -> global as 16-byte vectors (every piece starts on a 32-byte boundary) =====
        // [profiles/r2: as bulk (TMA) stores the kernel ran at the latency of cp.async.bulk.wait_group.read; as ONE warp copying
        // every chunk in order the warp's own ~350 clk per chunk (wake-up, LDS, arrive — serial) capped the kernel: removing the
        // ring took 0.74 of 1.83 ms off the skeleton.  Now four warps, each owning one slot of the ring: chunk oc -> warp oc % 4.]
        const int sw = warp - 12;
        uint32_t gt = 0;
        VS_FOR_EACH_TILE {
            const long long hand0 = (long long)tile * VS_NH;
            for (int t = 0; t < VS_NT; ++t, ++gt) {
                for (int ch = sw; ch < VS_NCH; ch += VS_OSTAGES) {
                    const uint32_t ob = sw;                               // chunk counter gt * 16 + ch = sw (mod 4)
                    if (variant & 0x2000) continue;                     // experiment: no result ring
                    vs_wait(&S.out_full[ob], (ch >> 2) & 1);            // the slot's use count = 4 gt + ch / 4
                    // carried floats of the previous vertex tile go in front of the results
                    if (t >= 1) {
                        const int hl = lane >> 3, i = lane & 7;
                        if (i < vs_d(hl)) S.out[ob][hl][i] = S.carry[gt & 1][ch * VS_HC + hl][i];
                    }
                    __syncwarp();
                    if (t >= 1 && t < VS_NT - 1) {
                        // 4 rows x 1536 B = 4 x 96 float4: three per lane per row, all loads first
                        float4 v[VS_HC][3];
#pragma unroll
                        for (int hl = 0; hl < VS_HC; ++hl) {
                            const float4* src = reinterpret_cast<const float4*>(&S.out[ob][hl][0]);
#pragma unroll
                            for (int k = 0; k < 3; ++k) v[hl][k] = src[lane + 32 * k];
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&S.out_empty[ob]));        // the rows are in registers
#pragma unroll
                        for (int hl = 0; hl < VS_HC; ++hl) {
                            const long long hand = hand0 + ch * VS_HC + hl;
                            if (hand < B && !(variant & 0x100)) {                     // 0x100: experiment, no global stores
                                float4* dst = reinterpret_cast<float4*>(verts + (size_t)hand * NVC + 3 * VS_M * t - vs_d(hl));
#pragma unroll
                                for (int k = 0; k < 3; ++k) __stcs(dst + lane + 32 * k, v[hl][k]);
                            }
                        }
                    } else {
                        // tile 0: the row's first floats up to its first 32-byte boundary are plain stores, the rest vectors;
                        // the short last tile (10 vertices + carry): plain stores
                        for (int hl = 0; hl < VS_HC; ++hl) {
                            const long long hand = hand0 + ch * VS_HC + hl;
                            if (hand >= B) break;
                            const int d = vs_d(hl);
                            float* grow = verts + (size_t)hand * NVC;
                            const float* row = S.out[ob][hl];
                            if (t == 0) {
                                const int a = (8 - d) & 7, sh = vs_shift(0, hl);
                                if (lane < a) grow[lane] = row[sh + lane];
                                const float4* src = reinterpret_cast<const float4*>(row + sh + a);
                                float4* dst = reinterpret_cast<float4*>(grow + a);
                                const int n16 = (3 * VS_M - d - a) / 4;                 // 94 or 96
                                for (int i = lane; i < n16; i += 32) __stcs(dst + i, src[i]);
                            } else {
                                const int n = NVC - (VS_NT - 1) * 3 * VS_M + d;          // 30 + d floats
                                for (int i = lane; i < n; i += 32) grow[(VS_NT - 1) * 3 * VS_M - d + i] = row[i];
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&S.out_empty[ob]));
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: thread = vertex; warps 4-7 take the even chunks of a tile, warps 8-11 the odd ones =====
        // [profiles/r2: with all eight warps on every chunk (two hands each) a chunk took ~1 500 clk — three mbarrier round
        // trips, five TMEM loads and their wait per 2 hands of work, serial in every warp — and the tensor pipe sat at 14 %.
        // Now a warp does all four hands of every other chunk: half the synchronisation per hand, and the other set's
        // round hides it.]
        const int q = warp & 3, set = (warp - 4) >> 2;
        const int vl = q * 32 + lane;                                  // vertex inside the tile
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const float osv = exp2f(-(float)(hdr->basis_scale_log2 + hdr->feat_scale_log2));
        const float ost = exp2f(-(float)(VS_W_SCALE_LOG2 + VS_BONE_SCALE_LOG2));
        uint32_t gt = 0, rounds = 0;                                   // rounds: chunks this set has taken (its transform stage's phase)
        VS_FOR_EACH_TILE {
            const long long hand0 = (long long)tile * VS_NH;
            for (int t = 0; t < VS_NT; ++t, ++gt) {
                const int vtx = t * VS_M + vl;
                const bool valid = vtx < NV;
                const float4 tm = vs_tmpl[vtx];
                int tipslot = -1;
#pragma unroll
                for (int i = 0; i < 5; ++i) if (vtx == c_vs_tip_vert[i]) tipslot = c_vs_tip_slot[i];
                const bool carries = valid && t + 1 < VS_NT && vl >= VS_M - 2;
                const uint32_t vp_addr = tmem + lane_addr + (gt & 1) * VS_VP_COLS;
                vs_wait(&S.vp_full[gt & 1], (gt >> 1) & 1);
                tc_fence_after();
                if (v_posed_t != nullptr) {
                    // rest-pose scratch for the skinning backward: v_posed_t[group][3 pos + p][32 hands], this warp's 32 hands
                    const long long group = (long long)tile * 2 + set;
                    const int pos3 = __float_as_int(tm.w);
                    const bool live = valid && pos3 >= 0 && group * 32 < B;
#pragma unroll 1
                    for (int p = 0; p < 3; ++p) {
                        uint32_t r[32];
                        tmem_ld32_nowait(vp_addr + p * VS_NH + set * 32, r);
                        tmem_ld_wait();
                        if (live) {
                            const float tp = p == 0 ? tm.x : (p == 1 ? tm.y : tm.z);
                            float4* dst = reinterpret_cast<float4*>(v_posed_t + ((size_t)group * SK_NCOORD + pos3 + p) * 32);
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                __stcs(dst + i, make_float4(fmaf(__uint_as_float(r[4 * i]), osv, tp), fmaf(__uint_as_float(r[4 * i + 1]), osv, tp),
                                                            fmaf(__uint_as_float(r[4 * i + 2]), osv, tp), fmaf(__uint_as_float(r[4 * i + 3]), osv, tp)));
                        }
                    }
                }
#pragma unroll 1
                for (int ch = set; ch < VS_NCH; ch += 2, ++rounds) {
                    vs_wait(&S.t_full[set], rounds & 1);
                    tc_fence_after();
                    const uint32_t t_addr = tmem + lane_addr + VS_T_COL0 + set * VS_TN;
                    uint32_t T[VS_TN], X[VS_HC], Y[VS_HC], Z[VS_HC];
                    const uint32_t x_addr = vp_addr + ch * VS_HC;
                    if (!(variant & 0x1000)) {                          // 0x1000: experiment, no TMEM loads
                        tmem_ld32_nowait(t_addr, T);
                        tmem_ld16_nowait(t_addr + 32, T + 32);
                        tmem_ld4_nowait(x_addr, X);
                        tmem_ld4_nowait(x_addr + VS_NH, Y);
                        tmem_ld4_nowait(x_addr + 2 * VS_NH, Z);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int i = 0; i < VS_TN; ++i) T[i] = i;
#pragma unroll
                        for (int i = 0; i < VS_HC; ++i) X[i] = Y[i] = Z[i] = i;
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&S.t_empty[set]));
                    // result rows of chunk oc = (chunks so far) -> ring slot; the set's slots alternate {set, set + 2}
                    const uint32_t oc = gt * VS_NCH + ch;
                    const uint32_t ob = oc % VS_OSTAGES;
                    if (variant & 0x2000) continue;                     // 0x2000: experiment, no result ring
                    vs_wait(&S.out_empty[ob], ((oc / VS_OSTAGES) & 1) ^ 1);
#pragma unroll
                    for (int hl = 0; hl < VS_HC; ++hl) {
                        if (variant & 0x200) break;                                 // 0x200: experiment, handshakes only
                        const float x = fmaf(__uint_as_float(X[hl]), osv, tm.x);
                        const float y = fmaf(__uint_as_float(Y[hl]), osv, tm.y);
                        const float z = fmaf(__uint_as_float(Z[hl]), osv, tm.z);
                        float o[3];
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const float* Ti = reinterpret_cast<const float*>(T) + hl * 12 + 4 * i;
                            o[i] = ost * fmaf(Ti[0], x, fmaf(Ti[1], y, fmaf(Ti[2], z, Ti[3])));
                        }
                        if (dbg != nullptr && tile == 0 && t == 0 && ch == 0) {
                            float* dd = dbg + ((size_t)hl * VS_M + vl) * 16;
#pragma unroll
                            for (int i = 0; i < 12; ++i) dd[i] = ost * __uint_as_float(T[hl * 12 + i]);
                            dd[12] = x; dd[13] = y; dd[14] = z; dd[15] = 0.f;
                        }
                        if (valid) {
                            float* row = S.out[ob][hl] + vs_shift(t, hl) + 3 * vl;
                            row[0] = o[0]; row[1] = o[1]; row[2] = o[2];
                        }
                        if (carries) {
                            // the last d floats of this tile's piece open the next tile's piece
                            const int d = vs_d(hl);
#pragma unroll
                            for (int i = 0; i < 3; ++i) {
                                const int j = 3 * vl + i - (3 * VS_M - d);
                                if (j >= 0) S.carry[(gt + 1) & 1][ch * VS_HC + hl][j] = o[i];
                            }
                        }
                        if (tipslot >= 0 && valid) {
                            const long long hand = hand0 + ch * VS_HC + hl;
                            if (hand < B) {
                                float* jo = joints + (size_t)hand * (NOUTJ * 3) + tipslot * 3;
                                jo[0] = o[0]; jo[1] = o[1]; jo[2] = o[2];
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&S.out_full[ob]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&S.vp_empty[gt & 1]));
            }
        }
    }
#undef VS_FOR_EACH_TILE
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                              // no CTA leaves while a peer may still signal its barriers
    if (warp == 2) tmem_dealloc(tmem, VS_TMEM_COLS);
}

}  // namespace

// ---------------------------------------------------------------- host side
size_t vskin_blob_bytes() { return vs_blob_layout().total; }

void vskin_pack(const float* basis, const float* skin_w, const int32_t* skin_b, const int32_t* sk_perm, int basis_scale_log2,
                void* host_section) {
    const VsBlobLayout L = vs_blob_layout();
    unsigned char* out = reinterpret_cast<unsigned char*>(host_section);
    memset(out, 0, L.total);
    const float sb = ldexpf(1.f, basis_scale_log2);
    // blend basis as the A operand: rows = vertices of the tile, one image per coordinate plane
    __half* bdst = reinterpret_cast<__half*>(out + L.basis);
    for (int t = 0; t < VS_NT; ++t)
        for (int p = 0; p < 3; ++p)
            for (int c = 0; c < TC_K_CHUNKS; ++c)
                for (int r = 0; r < VS_M; ++r)
                    for (int kk = 0; kk < TC_K_CHUNK; ++kk) {
                        const int v = t * VS_M + r, k = c * TC_K_CHUNK + kk;
                        const float x = (v < NV && k < TC_K_REAL) ? basis[(size_t)k * NVC + v * 3 + p] * sb : 0.f;
                        const __half hi = __float2half_rn(x);
                        const __half lo = __float2half_rn(x - __half2float(hi));
                        const size_t stage = (size_t)(t * 3 + p) * TC_K_CHUNKS + c;
                        const size_t in = (((size_t)(r >> 3) * (TC_K_CHUNK / 8) + (kk >> 3)) * 8 + (r & 7)) * 8 + (kk & 7);
                        bdst[stage * (VS_A_STAGE_BYTES / 2) + in] = hi;
                        bdst[stage * (VS_A_STAGE_BYTES / 2) + (VS_A_STAGE_BYTES / 4) + in] = lo;
                    }
    // dense skinning weights W[vertex][bone] in three fp16 splits, K-major (K = bone)
    std::vector<float> dense((size_t)NV * NJ, 0.f);
    for (int v = 0; v < NV; ++v)
        for (int s = 0; s < MAX_INFL; ++s) {
            const float w = skin_w[v * MAX_INFL + s];
            const int b = skin_b[v * MAX_INFL + s];
            if (w != 0.f && b >= 0 && b < NJ) dense[(size_t)v * NJ + b] += w;
        }
    __half* wdst = reinterpret_cast<__half*>(out + L.w);
    const float sw = (float)(1 << VS_W_SCALE_LOG2);
    for (int t = 0; t < VS_NT; ++t)
        for (int r = 0; r < VS_M; ++r)
            for (int k = 0; k < NJ; ++k) {
                const int v = t * VS_M + r;
                float x = v < NV ? dense[(size_t)v * NJ + k] * sw : 0.f;
                const size_t in = (((size_t)(r >> 3) * 2 + (k >> 3)) * 8 + (r & 7)) * 8 + (k & 7);
                for (int s = 0; s < VS_W_SPLITS; ++s) {
                    const __half h = __float2half_rn(x);
                    x -= __half2float(h);
                    wdst[((size_t)t * VS_W_SPLITS + s) * (VS_M * NJ) + in] = h;
                }
            }
    // v_template per vertex + the vertex' row in the block-order rest-pose scratch
    float* tm = reinterpret_cast<float*>(out + L.tmpl);
    std::vector<int> pos_of(NV, -1);
    for (int p = 0; p < SK_NPOS; ++p) if (sk_perm[p] >= 0 && sk_perm[p] < NV) pos_of[sk_perm[p]] = p;
    for (int v = 0; v < VS_NT * VS_M; ++v) {
        int pos3 = -1;
        for (int c = 0; c < 3; ++c) tm[v * 4 + c] = v < NV ? basis[(size_t)FEAT_ONE * NVC + v * 3 + c] : 0.f;
        if (v < NV && pos_of[v] >= 0) pos3 = pos_of[v] * 3;
        memcpy(&tm[v * 4 + 3], &pos3, sizeof(int));
    }
}

int launch_vskin_forward(const void* blob, const unsigned char* featp, const float* bone_t, int B, int mode,
                         float* verts, float* joints, float* v_posed_t, float* dbg, int variant, cudaStream_t s) {
    if (B <= 0) return 0;
    static SmemAttrOnce once;
    const size_t smem = sizeof(VsShared) + 128;
    if (int arc = ensure_dyn_smem(once, vskin_forward_kernel, smem)) return arc;
    const BlobLayout L = blob_layout();
    const unsigned char* tc = blob_ptr<unsigned char>(blob, L.total);
    const unsigned char* vs = tc + blend_tc_blob_bytes();
    const VsBlobLayout V = vs_blob_layout();
    const int ntiles = (B + VS_NH - 1) / VS_NH;
    if (const char* ev = getenv("MANO_B200_VSKIN_VARIANT")) variant |= (int)strtol(ev, nullptr, 0);   // experiments (profiles/r2)
    int t_products = (variant >> 4) & 7;
    if (t_products == 0) t_products = 4;
    // cluster size 1, 2 (default) or 4: the CTAs of a cluster share one multicast basis stream
    int csize = 2;
    if (const char* ev = getenv("MANO_B200_VSKIN_CLUSTER")) csize = atoi(ev);
    if (csize != 1 && csize != 2 && csize != 4) csize = 2;
    while (csize > 1 && ntiles < csize) csize >>= 1;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(VS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = NUM_SMS / csize;
    if (csize > 1) {                                                   // clusters that can be co-resident (GPC sizes strand a few SMs at 4)
        cfg.gridDim = dim3(NUM_SMS / csize * csize);
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, vskin_forward_kernel, &cfg) == cudaSuccess && n > 0 && n < max_clusters) max_clusters = n;
    }
    const int rounds = (ntiles + csize - 1) / csize;
    cfg.gridDim = dim3((unsigned)((rounds < max_clusters ? rounds : max_clusters) * csize));
    const TcBlobHeader* hdr = reinterpret_cast<const TcBlobHeader*>(tc);
    const unsigned char* basis_p = vs + V.basis;
    const unsigned char* w_p = vs + V.w;
    const float4* tmpl_p = reinterpret_cast<const float4*>(vs + V.tmpl);
    const int bp = mode == MB_MODE_F16X3 ? 3 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, vskin_forward_kernel, hdr, basis_p, w_p, tmpl_p, featp, bone_t, B, ntiles, bp, t_products,
                                       verts, joints, v_posed_t, dbg, variant);
    if (e != cudaSuccess) return (int)e;
    return cuda_rc();
}

}  // namespace mb
