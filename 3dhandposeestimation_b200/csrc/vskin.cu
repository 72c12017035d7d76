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
// Roles (384 threads, 1 CTA per SM, persistent over 64-hand tiles; a UNIT is one (hand tile, vertex tile) pair):
//   warp 0   blend issuer (one thread): the blend products of unit u + 1 run while the epilogue works on unit u
//   warp 1   transform issuer (one thread): T = W A of unit u, 4 hands per chunk
//   warp 2   TMEM allocation; basis producer (one thread): (tile, plane, K chunk) stages of 16 KB, always L2 hits, multicast
//            inside a cluster
//   warp 3   producer of the slower streams (one polling thread): the hand tile's feature rows (40 KB), the weight tiles (8 KB per
//            vertex tile, double buffered) and the transform products' fp16 x3 bone operand (72 KB per hand tile, written by
//            vs_bones_operand_kernel from the pose stage's fp32 transforms), chunk by chunk behind per-chunk barriers
//   warps 4-11  epilogue: warp % 4 = TMEM lane quarter (32 vertices); warps 4-7 own hands 0-31 of the tile, 8-11 hands 32-63
// TMEM (512 columns): ONE rest-position stage of 3 x 64 columns — every epilogue thread copies its 3 x 32 rest coordinates
// into registers at the start of a unit and frees the stage for the next unit's blend products at once — and a ring of
// six transform stages of 48 columns (three chunks in flight per epilogue warp set).
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
// 256-bit store (sm_100): eight consecutive floats = one full 32-byte sector, streaming
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

namespace {

constexpr int VS_THREADS = 384;                   // 12 warps: see the role list above
constexpr int VS_EPI_WARPS = 8;
constexpr int VS_ASTAGES = 5;                     // basis ring: 5 x 16 KB (6 measured the same)
constexpr int VS_TSTAGES = 3;                     // transform ring of 8-hand chunks: chunk counter mod 3; even counters -> warp set 0, odd -> set 1
constexpr int VS_HH = 4;                          // hands per epilogue register load of a chunk (48 columns)
constexpr int VS_HS = VS_NH / 2;                  // 32 hands per epilogue warp set
constexpr uint32_t VS_TMEM_COLS = 512;
constexpr uint32_t VS_VP_COLS = 3 * VS_NH;         // 192 columns: the rest positions of a unit
constexpr uint32_t VS_T_COL0 = VS_VP_COLS;         // 192: first transform column
constexpr uint32_t VS_W_COL0 = VS_T_COL0 + VS_TSTAGES * VS_TN;     // 480: the weight tiles (A operand of the transform products), two buffers of
constexpr uint32_t VS_W_COLS = VS_W_SPLITS * NJ / 2;               //      16 columns: [split][8 columns = 16 bones as fp16 pairs], lane = vertex
// instruction descriptors: f16 x f16 -> f32, M = 128
constexpr uint32_t VS_IDESC_BLEND = (1u << 4) | ((uint32_t)(VS_NH >> 3) << 17) | ((uint32_t)(VS_M >> 4) << 24);                // A, B K-major
constexpr uint32_t VS_IDESC_T = (1u << 4) | (1u << 16) | ((uint32_t)(VS_TN >> 3) << 17) | ((uint32_t)(VS_M >> 4) << 24);       // B MN-major

__constant__ int c_vs_tip_vert[5] = {333, 444, 672, 555, 745};
__constant__ int c_vs_tip_slot[5] = {4, 8, 12, 16, 20};

struct VsShared {
    alignas(128) unsigned char feat[TC_K_CHUNKS][2][VS_NH * TC_K_CHUNK * 2];      // 40 KB: [K chunk][hi, lo][64 hands x 32]
    alignas(128) unsigned char bones[VS_NCH][VS_BONE_SPLITS][VS_BONE_CHUNK_BYTES]; // 72 KB
    alignas(128) unsigned char a[VS_ASTAGES][VS_A_STAGE_BYTES];                    // 64 KB basis ring
    alignas(8) unsigned long long a_full[VS_ASTAGES], a_empty[VS_ASTAGES];
    unsigned long long w_full[2], w_empty[2];
    unsigned long long feat_full, feat_empty;
    unsigned long long bones_full[VS_NCH], bones_empty[VS_NCH];
    unsigned long long vp_full, vp_empty;
    unsigned long long t_full[VS_TSTAGES], t_empty[VS_TSTAGES];
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

// -DVS_PROFILE: every role of CTA 0 accumulates the cycles it spends in each kind of wait and its total loop time into
// dbg[8192 + 8 * role ...] (profiles/tools/vskin_roles.py); compiled out otherwise
#ifdef VS_PROFILE
#define VS_WAIT(bar, parity, slot) do { const long long _t0 = clock64(); vs_wait(bar, parity); prof[slot] += clock64() - _t0; } while (0)
#define VS_PROF_DECL long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long prof_t0 = clock64()
#define VS_PROF_STORE(role) do { if (dbg != nullptr && blockIdx.x == 0) { prof[7] = clock64() - prof_t0; \
    for (int _i = 0; _i < 8; ++_i) dbg[8192 + 8 * (role) + _i] = (float)prof[_i]; } } while (0)
#define VS_TIC const long long _tic = clock64()
#define VS_TOC(slot) prof[slot] += clock64() - _tic
#else
#define VS_TIC do {} while (0)
#define VS_TOC(slot) do {} while (0)
#define VS_WAIT(bar, parity, slot) vs_wait(bar, parity)
#define VS_PROF_DECL do {} while (0)
#define VS_PROF_STORE(role) do {} while (0)
#endif

// the j-th transform chunk ISSUED of a unit is chunk (j >> 1) + 8 (j & 1): hands 4 ch .. 4 ch + 3 of the tile, so that the two
// epilogue warp sets (hands 0-31, 32-63) are served alternately
__device__ __forceinline__ int vs_chunk_of(int j) { return (j >> 1) + (VS_NCH / 2) * (j & 1); }

// ---- bone operand images: fp32 transforms bone_t[group][bone][hand % 32][12] -> fp16 x3 MN-major core matrices, per 8-hand chunk
// [hand tile][chunk 8][split 3][3072 B: 12 n-groups x 2 k-groups x 8 k x 8 n].  The 12 elements of 8 consecutive hands of one bone
// are 96 contiguous floats = one K row (k = bone) of a chunk's B operand, n = (hand % 8) * 12 + element: a lane converts 8 of
// them (one 16-byte n-group) into the three splits a = a1 + a2 + a3 of 2^4 a; one warp per chunk.
__global__ void __launch_bounds__(256)
vs_bones_operand_kernel(const float* __restrict__ bone_t, int B, long long nchunks, unsigned char* __restrict__ bones_op) {
    const int lane = threadIdx.x & 31;
    const long long ngroups = ((long long)B + 31) >> 5;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long cidx = warp0; cidx < nchunks; cidx += nwarps) {
        const long long group = cidx / (32 / VS_HC);                    // 4 chunks of 8 hands per 32-hand group
        const float* src0 = bone_t + (size_t)group * (NJ * BONE_F * 32) + VS_TN * (int)(cidx % (32 / VS_HC));
        unsigned char* out = bones_op + (size_t)cidx * (VS_BONE_SPLITS * VS_BONE_CHUNK_BYTES);
        constexpr int NR = NJ * (VS_TN / 8) / 32;                       // 192 (bone, n-group) items per chunk: six per lane
        float4 v[NR][2];
#pragma unroll
        for (int r = 0; r < NR; ++r) {                                  // all twelve loads of the lane in flight before the first conversion
            const int i = lane + 32 * r, k = i / (VS_TN / 8), g = i - (VS_TN / 8) * k;
            v[r][0] = make_float4(0.f, 0.f, 0.f, 0.f);
            v[r][1] = v[r][0];
            if (group < ngroups) {
                const float4* p = reinterpret_cast<const float4*>(src0 + (size_t)k * (BONE_F * 32) + 8 * g);
                v[r][0] = __ldcs(p); v[r][1] = __ldcs(p + 1);
            }
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int i = lane + 32 * r, k = i / (VS_TN / 8), g = i - (VS_TN / 8) * k;
            float x[8] = {v[r][0].x, v[r][0].y, v[r][0].z, v[r][0].w, v[r][1].x, v[r][1].y, v[r][1].z, v[r][1].w};
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] *= (float)(1 << VS_BONE_SCALE_LOG2);
            unsigned char* dst = out + g * 256 + (k >> 3) * 128 + (k & 7) * 16;
#pragma unroll
            for (int sp = 0; sp < VS_BONE_SPLITS; ++sp) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const __half2 h = __floats2half2_rn(x[2 * e], x[2 * e + 1]);
                    const float2 f = __half22float2(h);
                    x[2 * e] -= f.x;
                    x[2 * e + 1] -= f.y;
                    pk[e] = *reinterpret_cast<const uint32_t*>(&h);
                }
                __stcs(reinterpret_cast<uint4*>(dst + sp * VS_BONE_CHUNK_BYTES), make_uint4(pk[0], pk[1], pk[2], pk[3]));
            }
        }
    }
}

__global__ void __launch_bounds__(VS_THREADS, 1)
vskin_forward_kernel(const TcBlobHeader* __restrict__ hdr, const unsigned char* __restrict__ vs_basis,
                     const unsigned char* __restrict__ vs_w, const float4* __restrict__ vs_tmpl,
                     const unsigned char* __restrict__ featp, const unsigned char* __restrict__ bones_op,
                     int B, int ntiles, int blend_products, int t_products,
                     float* __restrict__ verts, float* __restrict__ joints, float* __restrict__ v_posed_t,
                     float* __restrict__ dbg, int variant) {
    // [profiles/r2: rounding the base up through uintptr_t hides the shared address space from the compiler — every access of
    // this kernel is a generic LD / ST.  Declared `__align__(1024)` and used directly (LDS / STS, as blend_tc.cu now does) this
    // kernel measured SLOWER (1.78 against 1.63 ms per 262 144 hands, twice), so the generic form stays here.]
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
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&S.w_full[s]), VS_EPI_WARPS); mbar_init(smem_u32(&S.w_empty[s]), 1); }
        mbar_init(smem_u32(&S.vp_full), 1); mbar_init(smem_u32(&S.vp_empty), VS_EPI_WARPS);
        for (int s = 0; s < VS_TSTAGES; ++s) { mbar_init(smem_u32(&S.t_full[s]), 1); mbar_init(smem_u32(&S.t_empty[s]), VS_EPI_WARPS / 2); }
        mbar_init(smem_u32(&S.feat_full), 1); mbar_init(smem_u32(&S.feat_empty), 1);
        for (int s = 0; s < VS_NCH; ++s) { mbar_init(smem_u32(&S.bones_full[s]), 1); mbar_init(smem_u32(&S.bones_empty[s]), 1); }
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
            // ===== blend issuer (one thread) =====
            // Per (plane, K chunk) stage of the basis ring: wait for its bytes, issue its 6 MMAs, commit -> a_empty.
            // [profiles/r2: ONE thread polling both the blend and the transform cursor spent ~400 clk per event — 31 events per
            // unit = 13 000 clk, the whole kernel time with every MMA removed: two issuing threads with blocking waits.  With the
            // basis refills in this thread too it was busy 9 000 clk per unit (the critical chain of the kernel: 15 400 per unit):
            // the refills moved to the producer thread of warp 2, which idled 98 % of the time.]
            constexpr uint32_t NS = VS_ASTAGES;
            const uint32_t hi_k512 = desc_hi(TC_SBO);
            const uint32_t a_lo0 = desc_lo(smem_u32(S.a[0]), TC_LBO), f_lo0 = desc_lo(smem_u32(S.feat[0][0]), TC_LBO);
            int n_it = 0;
            VS_FOR_EACH_TILE { (void)tile; ++n_it; }
            VS_PROF_DECL;
            uint32_t st = 0;                                            // stage counter over the whole kernel
            uint32_t g = 0;                                             // unit counter
            for (int it = 0; it < n_it; ++it) {
                VS_WAIT(&S.feat_full, it & 1, 0);
                for (int t = 0; t < VS_NT; ++t, ++g) {
                    VS_WAIT(&S.vp_empty, (g & 1) ^ 1, 1);                  // the epilogue has copied the previous unit's rest positions out
                    tc_fence_after();
#pragma unroll 1
                    for (int pl = 0; pl < 3; ++pl) {
                        const uint32_t d = tmem + pl * VS_NH;
#pragma unroll
                        for (int c = 0; c < TC_K_CHUNKS; ++c, ++st) {
                            const uint32_t slot = st % NS;
                            VS_WAIT(&S.a_full[slot], (st / NS) & 1, 2);
                            tc_fence_after();
                            const uint32_t a_st = a_lo0 + ((slot * VS_A_STAGE_BYTES) >> 4);
                            const uint32_t f_st = f_lo0 + ((c * 2 * 4096) >> 4);
                            if (!(variant & 0x400)) {                   // 0x400: experiment, no blend products
#pragma unroll
                                for (int j = 0; j < TC_K_CHUNK / 16; ++j) {
                                    const uint32_t a_hi = a_st + ((j * 2 * (int)TC_LBO) >> 4), a_lo = a_hi + ((VS_A_STAGE_BYTES / 2) >> 4);
                                    const uint32_t b_hi = f_st + ((j * 2 * (int)TC_LBO) >> 4), b_lo = b_hi + (4096 >> 4);
                                    if (j == 0 && c == 0) umma_f16_lohi<false>(d, a_hi, hi_k512, b_hi, hi_k512, VS_IDESC_BLEND);
                                    else umma_f16_lohi<true>(d, a_hi, hi_k512, b_hi, hi_k512, VS_IDESC_BLEND);
                                    if (blend_products == 3) {
                                        umma_f16_lohi<true>(d, a_lo, hi_k512, b_hi, hi_k512, VS_IDESC_BLEND);
                                        umma_f16_lohi<true>(d, a_hi, hi_k512, b_lo, hi_k512, VS_IDESC_BLEND);
                                    }
                                }
                            }
                            if (variant & 0x10000) mbar_arrive(smem_u32(&S.a_empty[slot]));      // 0x10000: experiment (with 0x400, csize 1): plain arrive
                            else if (csize == 1) tc_commit(smem_u32(&S.a_empty[slot]));
                            else tc_commit_multicast(smem_u32(&S.a_empty[slot]), cmask);    // every CTA of the cluster refills this slot
                        }
                    }
                    tc_commit(smem_u32(&S.vp_full));
                }
                tc_commit(smem_u32(&S.feat_empty));
            }
            VS_PROF_STORE(0);
        }
    } else if (warp == 2) {
        if (elect_one()) {
            // ===== basis producer (one thread, blocking waits: this stream is the one the blend issuer waits for) =====
            // The basis is the same 105-stage stream for every hand tile, kept in L2 (evict-last) and, inside a cluster, loaded
            // once and multicast.
            const uint64_t keep = l2_policy_evict_last();
            const uint32_t slice = VS_A_STAGE_BYTES / csize;           // this CTA's share of every basis stage
            constexpr uint32_t NS = VS_ASTAGES;
            int n_it = 0;
            VS_FOR_EACH_TILE { (void)tile; ++n_it; }
            const uint32_t total_a = (uint32_t)n_it * (VS_NT * VS_STAGES_PER_TILE);
            VS_PROF_DECL;
            uint32_t i = 0;                                             // stage inside the 105-stage stream
            for (uint32_t na = 0; na < total_a; ++na) {
                const uint32_t slot = na % NS;
                // stage na -> slot na % NS: free once the MMAs of stage na - NS are done
                if (na >= NS) VS_WAIT(&S.a_empty[slot], ((na / NS) & 1) ^ 1, 0);
                if (variant & 0x8000) mbar_arrive(smem_u32(&S.a_full[slot]));              // 0x8000: experiment, no basis bytes
                else {
                    mbar_expect_tx(smem_u32(&S.a_full[slot]), VS_A_STAGE_BYTES);
                    if (csize == 1)
                        bulk_g2s_hint(smem_u32(S.a[slot]), vs_basis + (size_t)i * VS_A_STAGE_BYTES, VS_A_STAGE_BYTES, smem_u32(&S.a_full[slot]), keep);
                    else
                        bulk_g2s_multicast(smem_u32(S.a[slot]) + crank * slice, vs_basis + (size_t)i * VS_A_STAGE_BYTES + crank * slice, slice,
                                           smem_u32(&S.a_full[slot]), cmask, keep);
                }
                if (++i == VS_NT * VS_STAGES_PER_TILE) i = 0;
            }
            VS_PROF_STORE(2);
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ===== transform issuer: T = W A, 4 hands per chunk, through the six-stage ring =====
            // [profiles/r2: issued by `lane == 0` the compiler wrapped every tcgen05.mma in a divergence "waterfall" (ELECT /
            // R2UR.BROADCAST / BRA.U.ANY) and rebuilt 64-bit descriptors per instruction: elect.sync and 32-bit descriptor halves.]
            // bones: MN-major, n-groups 256 B apart (SBO), k-groups 128 B apart (LBO)
            const uint32_t b_lo0 = desc_lo(smem_u32(&S.bones[0][0][0]), 128);
            const uint32_t hi_b = desc_hi(256);
            uint32_t g = 0, stage = 0, phase = 0, it = 0;
            VS_PROF_DECL;
            VS_FOR_EACH_TILE {
                (void)tile;
                for (int t = 0; t < VS_NT; ++t, ++g) {
                    VS_WAIT(&S.w_full[g & 1], (g >> 1) & 1, 0);
                    // the unit's weight tile sits in TMEM (written by the epilogue warps): buffer g & 1, splits 8 columns apart
                    const uint32_t w1 = tmem + VS_W_COL0 + (g & 1) * VS_W_COLS, w2 = w1 + NJ / 2;
#pragma unroll 1
                    for (int j = 0; j < VS_NCH; ++j) {
                        const int ch = vs_chunk_of(j);
                        if (t == 0) VS_WAIT(&S.bones_full[ch], it & 1, 1);            // the tile's first unit waits for the converted bones
                        VS_WAIT(&S.t_empty[stage], phase ^ 1, 2);
                        tc_fence_after();
                        const uint32_t d = tmem + VS_T_COL0 + stage * VS_TN;
                        const uint32_t a1 = b_lo0 + ((ch * VS_BONE_SPLITS * VS_BONE_CHUNK_BYTES) >> 4);
                        const uint32_t a2 = a1 + (VS_BONE_CHUNK_BYTES >> 4), a3 = a2 + (VS_BONE_CHUNK_BYTES >> 4);
                        // smallest products first, w1 a1 last: the tensor core TRUNCATES its fp32 accumulator at the running
                        // sum's magnitude on every MMA; with the corrections accumulated among themselves first, only the
                        // last addition rounds at full magnitude
                        if (!(variant & 0x800)) {                       // 0x800: experiment, no transform products
                            if (t_products >= 4) {
                                umma_f16_ts<false>(d, w1, a3, hi_b, VS_IDESC_T);
                                if (t_products >= 5) umma_f16_ts<true>(d, w2, a2, hi_b, VS_IDESC_T);
                                umma_f16_ts<true>(d, w1, a2, hi_b, VS_IDESC_T);
                                umma_f16_ts<true>(d, w2, a1, hi_b, VS_IDESC_T);
                                umma_f16_ts<true>(d, w1, a1, hi_b, VS_IDESC_T);
                            } else if (t_products == 3) {
                                umma_f16_ts<false>(d, w1, a2, hi_b, VS_IDESC_T);
                                umma_f16_ts<true>(d, w2, a1, hi_b, VS_IDESC_T);
                                umma_f16_ts<true>(d, w1, a1, hi_b, VS_IDESC_T);
                            } else {
                                umma_f16_ts<false>(d, w1, a1, hi_b, VS_IDESC_T);
                            }
                        }
                        if (variant & 0x20000) mbar_arrive(smem_u32(&S.t_full[stage]));          // 0x20000: experiment (with 0x800): plain arrive
                        else tc_commit(smem_u32(&S.t_full[stage]));
                        if (t == VS_NT - 1) tc_commit(smem_u32(&S.bones_empty[ch]));        // the tile's last use of the chunk's bones
                        if (++stage == VS_TSTAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(smem_u32(&S.w_empty[g & 1]));                 // the unit's products are done with its weight tile
                }
                ++it;
            }
            VS_PROF_STORE(1);
        }
    } else if (warp == 3) {
        if (elect_one()) {
            // ===== producer of the slower streams (one polling thread): feature rows of a hand tile, weight tiles, bone operand =====
            // Bone operand: the chunk's three fp16 splits (4.5 KB, written by vs_bones_operand_kernel) as one bulk copy; chunks are
            // handed over one by one (the previous tile's last unit frees them one by one), in the order the transform issuer
            // takes them.  [profiles/r2: converted HERE from the pose stage's fp32 transforms by this warp alone, a chunk took
            // 2 500 clk (fp32 <-> fp16 conversions run at a quarter of the fp32 rate) and the first unit of every hand tile waited
            // for all sixteen: 40 000 of a tile's 97 000 clk.]
            const uint64_t once = l2_policy_evict_first();
            int n_it = 0;
            VS_FOR_EACH_TILE { (void)tile; ++n_it; }
            const uint32_t total_b = (uint32_t)n_it * VS_NCH;
            uint32_t nb = 0;                                           // bone chunks requested so far
            int fit = 0, rnd_f = cluster_id, rnd_b = cluster_id;        // feature tiles requested; rounds of the next feature / bone tile
            VS_PROF_DECL;
            long long idle_since = -1;
            while (nb < total_b || fit < n_it) {
                bool progressed = false;
                // bone chunk nb: tile iteration nb / 16, issue index nb % 16
                if (nb < total_b) {
                    const uint32_t bit = nb / VS_NCH;
                    const int ch = vs_chunk_of((int)(nb % VS_NCH));
                    if (bit == 0 || mbar_test_wait(smem_u32(&S.bones_empty[ch]), (bit & 1) ^ 1)) {
                        const int tile = rnd_b * (int)csize + (int)crank;
                        const int ltile = tile < ntiles ? tile : ntiles - 1;
                        mbar_expect_tx(smem_u32(&S.bones_full[ch]), VS_BONE_SPLITS * VS_BONE_CHUNK_BYTES);
                        bulk_g2s_hint(smem_u32(&S.bones[ch][0][0]),
                                      bones_op + (size_t)ltile * VS_BONE_TILE_BYTES + (size_t)ch * (VS_BONE_SPLITS * VS_BONE_CHUNK_BYTES),
                                      VS_BONE_SPLITS * VS_BONE_CHUNK_BYTES, smem_u32(&S.bones_full[ch]), once);
                        if (++nb % VS_NCH == 0) rnd_b += nclusters;
                        progressed = true;
                    }
                }
                // feature rows of the next hand tile
                if (fit < n_it && (fit == 0 || mbar_test_wait(smem_u32(&S.feat_empty), (fit & 1) ^ 1))) {
                    const int tile = rnd_f * (int)csize + (int)crank;
                    const int ltile = tile < ntiles ? tile : ntiles - 1;
                    mbar_expect_tx(smem_u32(&S.feat_full), TC_K_CHUNKS * 2 * 4096);
                    const unsigned char* fsrc = featp + (size_t)(ltile >> 1) * TC_A_TILE_BYTES + (size_t)(ltile & 1) * 4096;
                    for (int c = 0; c < TC_K_CHUNKS; ++c)
                        for (int sp = 0; sp < 2; ++sp)
                            bulk_g2s_hint(smem_u32(S.feat[c][sp]), fsrc + (size_t)c * TC_A_STAGE_BYTES + (size_t)sp * TC_A_BLOCK_BYTES, 4096,
                                          smem_u32(&S.feat_full), once);
                    ++fit;
                    rnd_f += nclusters;
                    progressed = true;
                }
                if (progressed) idle_since = -1;
                else if (idle_since < 0) idle_since = clock64();
                else if (clock64() - idle_since > 2000000000LL) __trap();
            }
            VS_PROF_STORE(3);
        }
    } else if (warp >= 4) {
        // ===== epilogue: thread = vertex; the warp set owns 32 hands of the tile, whose rest coordinates it keeps in registers =====
        // [profiles/r2: (a) two rest-position stages + two transform stages in TMEM: a warp set had ONE chunk in flight, and the
        // scheduler -> tensor pipe -> epilogue -> scheduler round trip (~1 000 clk) was exposed on every chunk; (b) the result rows
        // went through a 4-slot ring to four store warps: +300 clk per chunk of handshakes.  Now the rest positions leave TMEM at
        // once (96 registers), which buys six transform stages, and a thread stores its own results.]
        const int q = warp & 3, set = (warp - 4) >> 2;
        const int vl = q * 32 + lane;                                  // vertex inside the tile
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const float osv = exp2f(-(float)(hdr->basis_scale_log2 + hdr->feat_scale_log2));
        const float ost = exp2f(-(float)(VS_W_SCALE_LOG2 + VS_BONE_SCALE_LOG2));
        uint32_t g = 0;                                                // unit counter
        int n_units = 0;
        VS_FOR_EACH_TILE { (void)tile; n_units += VS_NT; }
        // The weight tile of a unit (A operand of its transform products) lives in TENSOR MEMORY, lane = vertex: every epilogue
        // thread stores its own vertex' 16 weights of one fp16 split (set 0: hi, set 1: lo; 8 columns) one unit ahead.
        // [profiles/r2: as a shared-memory operand the 4 KB tile was re-read by each of the 64 products of a unit — 256 KB of the
        // 890 KB per unit that the shared-memory pipe (75 % busy) serves to the tensor core.]
        auto w_row_load = [&](int vt, uint4 (&wr)[2]) {
            const uint4* p = reinterpret_cast<const uint4*>(vs_w + ((size_t)(vt * VS_M + vl) * VS_W_SPLITS + set) * (NJ * 2));
            wr[0] = __ldg(p); wr[1] = __ldg(p + 1);
        };
        auto w_row_store = [&](uint32_t unit, const uint4 (&wr)[2]) {
            if (unit >= 2) vs_wait(&S.w_empty[unit & 1], ((unit >> 1) & 1) ^ 1);   // the products of unit - 2 are done with the buffer
            tc_fence_after();
            const uint32_t r[8] = {wr[0].x, wr[0].y, wr[0].z, wr[0].w, wr[1].x, wr[1].y, wr[1].z, wr[1].w};
            tmem_st8(tmem + lane_addr + VS_W_COL0 + (unit & 1) * VS_W_COLS + set * (NJ / 2), r);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&S.w_full[unit & 1]));
        };
        if (n_units > 0) { uint4 wr[2]; w_row_load(0, wr); w_row_store(0, wr); }
        uint32_t tcnt = set;                                           // this set's next chunk counter (issue order, all units)
        VS_PROF_DECL;
        VS_FOR_EACH_TILE {
            const long long hand0 = (long long)tile * VS_NH + set * VS_HS;
            for (int t = 0; t < VS_NT; ++t, ++g) {
                const int vtx = t * VS_M + vl;
                const bool valid = vtx < NV;
                const float4 tm = vs_tmpl[vtx];
                const bool has_next = (int)g + 1 < n_units;
                int tipslot = -1;
#pragma unroll
                for (int i = 0; i < 5; ++i) if (vtx == c_vs_tip_vert[i]) tipslot = c_vs_tip_slot[i];
                // ---- rest positions of this thread's vertex for the set's 32 hands: TMEM -> registers, stage released
                float X[VS_HS], Y[VS_HS], Z[VS_HS];
                {
                    uint32_t rx[VS_HS], ry[VS_HS], rz[VS_HS];
                    VS_WAIT(&S.vp_full, g & 1, 0);
                    tc_fence_after();
                    VS_TIC;
                    const uint32_t vp_addr = tmem + lane_addr + set * VS_HS;
                    tmem_ld32_nowait(vp_addr, rx);
                    tmem_ld32_nowait(vp_addr + VS_NH, ry);
                    tmem_ld32_nowait(vp_addr + 2 * VS_NH, rz);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&S.vp_empty));
#pragma unroll
                    for (int i = 0; i < VS_HS; ++i) {
                        X[i] = fmaf(__uint_as_float(rx[i]), osv, tm.x);
                        Y[i] = fmaf(__uint_as_float(ry[i]), osv, tm.y);
                        Z[i] = fmaf(__uint_as_float(rz[i]), osv, tm.z);
                    }
                    VS_TOC(5);
                }
                if (has_next) { uint4 wnext[2]; w_row_load(t + 1 < VS_NT ? t + 1 : 0, wnext); w_row_store(g + 1, wnext); }
                if (v_posed_t != nullptr) {
                    // rest-pose scratch for the skinning backward: v_posed_t[group][3 pos + p][32 hands], this set's 32 hands = one
                    // hand group.  A thread holds the 32 hands of ITS vertex and the vertex' three scratch rows are 384 contiguous
                    // bytes: twelve 256-bit stores (sm_100: one full 32-byte sector each) straight from the registers.
                    // [profiles/r2, scratch cost per 2^20 hands: 16-byte stores from the registers (half sectors) 4.4 ms; transposed
                    // through a per-warp shared-memory tile with 4-byte accesses 3.1 ms, with 16-byte accesses 2.1 ms; this 1.9 ms]
                    const long long group = (long long)tile * 2 + set;
                    const int pos3 = valid ? __float_as_int(tm.w) : -1;
                    if (group * 32 < B && pos3 >= 0) {
                        // 0x40000: experiment — every group's scratch lands on group 0's 300 KB (L2-resident: the stores' SM-side cost
                        // without their DRAM traffic); 0x80000: experiment — no scratch stores at all
                        float* dst = v_posed_t + ((size_t)((variant & 0x40000) ? 0 : group) * SK_NCOORD + pos3) * 32;
                        if (!(variant & (0x80000 | 0x100000))) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                st_global_v8(dst + 8 * k, &X[8 * k]);
                                st_global_v8(dst + 32 + 8 * k, &Y[8 * k]);
                                st_global_v8(dst + 64 + 8 * k, &Z[8 * k]);
                            }
                        }
                    }
                }
                // this warp's run of a hand's row: floats [3 (128 t + 32 q), + 96) of verts[hand]; 16-byte aligned for even hands
                const int nfl = min(96, max(0, (NV - (t * VS_M + q * 32)) * 3));           // 96, or 30 / 0 in the last tile
                float* run0 = verts + (size_t)hand0 * NVC + 3 * (t * VS_M + q * 32);
#pragma unroll
                for (int c8 = 0; c8 < VS_NCH / 2; ++c8, tcnt += 2) {
                    const uint32_t stage = tcnt % VS_TSTAGES, use = tcnt / VS_TSTAGES;
                    VS_WAIT(&S.t_full[stage], use & 1, 1);
                    tc_fence_after();
                    const uint32_t t_addr = tmem + lane_addr + VS_T_COL0 + stage * VS_TN;
#pragma unroll
                  for (int half = 0; half < VS_HC / VS_HH; ++half) {   // the chunk's 96 columns in two register loads of 4 hands
                    VS_TIC;
                    uint32_t T[VS_HH * BONE_F];
                    tmem_ld32_nowait(t_addr + half * (VS_HH * BONE_F), T);
                    tmem_ld16_nowait(t_addr + half * (VS_HH * BONE_F) + 32, T + 32);
                    tmem_ld_wait();
                    if (half == VS_HC / VS_HH - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&S.t_empty[stage]));
                    }
                    VS_TOC(2);
                    if (variant & 0x200) continue;                     // 0x200: experiment, handshakes only
                    long long _tic2 = 0;
#ifdef VS_PROFILE
                    _tic2 = clock64();
#endif
#pragma unroll
                    for (int hl = 0; hl < VS_HH; ++hl) {
                        const int hi = c8 * VS_HC + half * VS_HH + hl; // hand inside the set
                        const float x = X[hi], y = Y[hi], z = Z[hi];
                        float o[3];
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const float* Ti = reinterpret_cast<const float*>(T) + hl * 12 + 4 * i;
                            o[i] = ost * fmaf(Ti[0], x, fmaf(Ti[1], y, fmaf(Ti[2], z, Ti[3])));
                        }
                        if (dbg != nullptr && tile == 0 && t == 0 && set == 0 && c8 == 0 && half == 0) {
                            float* dd = dbg + ((size_t)hl * VS_M + vl) * 16;
#pragma unroll
                            for (int i = 0; i < 12; ++i) dd[i] = ost * __uint_as_float(T[hl * 12 + i]);
                            dd[12] = x; dd[13] = y; dd[14] = z; dd[15] = 0.f;
                        }
                        // results leave straight from the registers: 3 x 4-byte stores per hand, a warp instruction covering a
                        // 384-byte run at a 12-byte lane stride — the three of them complete every 32-byte sector back to back
                        // in L2.  [profiles/r2: staged through shared memory into 16-byte vector stores (3 STS + LDS.128 + STG.128
                        // per hand, 57 M of the kernel's 170 M shared-memory wavefronts) the kernel took 1.63 instead of 1.53 ms
                        // per 262 144 hands: the shared-memory pipe, 75 % busy, is what the MMAs' operand reads wait for.]
                        {
                            const long long hand = hand0 + hi;
                            if (hand < B && 3 * lane < nfl && !(variant & 0x100)) {      // 0x100: experiment, no global stores
                                float* dst = run0 + (size_t)hi * NVC + 3 * lane;
                                __stcs(dst, o[0]); __stcs(dst + 1, o[1]); __stcs(dst + 2, o[2]);
                                if ((variant & 0x100000) && v_posed_t != nullptr) {   // 0x100000: experiment, scratch in verts' layout
                                    float* ds = v_posed_t + (size_t)hand * NVC + 3 * (t * VS_M + q * 32) + 3 * lane;
                                    __stcs(ds, x); __stcs(ds + 1, y); __stcs(ds + 2, z);
                                }
                            }
                        }
                        if (tipslot >= 0 && valid) {
                            const long long hand = hand0 + hi;
                            if (hand < B) {
                                float* jo = joints + (size_t)hand * (NOUTJ * 3) + tipslot * 3;
                                jo[0] = o[0]; jo[1] = o[1]; jo[2] = o[2];
                            }
                        }
                    }
#ifdef VS_PROFILE
                    prof[3] += clock64() - _tic2;
#endif
                    (void)_tic2;
                  }
                }
            }
        }
        if (lane == 0) VS_PROF_STORE(warp);
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
    // dense skinning weights W[vertex][bone] as two fp16 splits, one 32-byte row per (vertex, split)
    std::vector<float> dense((size_t)NV * NJ, 0.f);
    for (int v = 0; v < NV; ++v)
        for (int s = 0; s < MAX_INFL; ++s) {
            const float w = skin_w[v * MAX_INFL + s];
            const int b = skin_b[v * MAX_INFL + s];
            if (w != 0.f && b >= 0 && b < NJ) dense[(size_t)v * NJ + b] += w;
        }
    __half* wdst = reinterpret_cast<__half*>(out + L.w);
    const float sw = (float)(1 << VS_W_SCALE_LOG2);
    for (int v = 0; v < VS_NT * VS_M; ++v)
        for (int k = 0; k < NJ; ++k) {
            float x = v < NV ? dense[(size_t)v * NJ + k] * sw : 0.f;
            for (int s = 0; s < VS_W_SPLITS; ++s) {
                const __half h = __float2half_rn(x);
                x -= __half2float(h);
                wdst[((size_t)v * VS_W_SPLITS + s) * NJ + k] = h;        // row of vertex v, split s: 16 bones, K = bone
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

int launch_vskin_forward(const void* blob, const unsigned char* featp, const float* bone_t, unsigned char* bones_op, int B, int mode,
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
        // asked once per (device, cluster size): the occupancy query is a slow host call — per launch it cost the pipelined
        // end-to-end arm ~0.7 ms of host time per chunk, 3 ms per step [profiles/r2]
        static int cached[16][5] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        int& slot = cached[dev & 15][csize];
        if (slot == 0) {
            cfg.gridDim = dim3(NUM_SMS / csize * csize);
            int n = 0;
            slot = (cudaOccupancyMaxActiveClusters(&n, vskin_forward_kernel, &cfg) == cudaSuccess && n > 0 && n < max_clusters) ? n : max_clusters;
        }
        max_clusters = slot;
    }
    const int rounds = (ntiles + csize - 1) / csize;
    cfg.gridDim = dim3((unsigned)((rounds < max_clusters ? rounds : max_clusters) * csize));
    const TcBlobHeader* hdr = reinterpret_cast<const TcBlobHeader*>(tc);
    const unsigned char* basis_p = vs + V.basis;
    const unsigned char* w_p = vs + V.w;
    const float4* tmpl_p = reinterpret_cast<const float4*>(vs + V.tmpl);
    const int bp = mode == MB_MODE_F16X3 ? 3 : 1;
    {
        // the transform products' B operand: every 4-hand chunk of every hand tile (chunks past the batch are zero)
        const long long nchunks = (long long)ntiles * VS_NCH;
        const long long blocks = (nchunks + 7) / 8;
        vs_bones_operand_kernel<<<(unsigned)(blocks < 8 * NUM_SMS ? blocks : 8 * NUM_SMS), 256, 0, s>>>(bone_t, B, nchunks, bones_op);
        if (int rc = cuda_rc()) return rc;
    }
    const unsigned char* bones_op_c = bones_op;
    cudaError_t e = cudaLaunchKernelEx(&cfg, vskin_forward_kernel, hdr, basis_p, w_p, tmpl_p, featp, bones_op_c, B, ntiles, bp, t_products,
                                       verts, joints, v_posed_t, dbg, variant);
    if (e != cudaSuccess) return (int)e;
    return cuda_rc();
}

}  // namespace mb
