// skin.cu — linear blend skinning over 778 vertices x 16 bones, forward and backward,
// register-blocked with LANE = HAND.
//
// Reference: MANOLayer.py:177-185 (T = sum_k w_vk A_k ; v' = T [v_posed;1]), :190-202 (the five
// fingertip vertices become joints 4,8,12,16,20) and :188,:204-205 (global rotation — already
// folded into the bone transforms by the pose stage, so it costs nothing here).
//
// Round-1 ncu of the first lane = hand kernel (bones in shared memory, one vertex at a time): l1tex data pipe 76 % busy,
// 1 321 shared-memory wavefronts per hand — three 16-byte bone rows per (vertex, bone) pair per lane.  This version keeps the
// bone transform in REGISTERS and amortises it over a block of 8 vertices (DESIGN.md 3.3 has the measured path):
//   * a warp owns 32 hands (lane = hand) and sweeps all vertices on its own — no block barriers;
//   * the rest-pose vertices arrive hand-minor (v_posed_t[group][coord][32], written that way by the blend GEMM's
//     epilogue) through a 2-slot bulk-copy (TMA engine) ring: a block of 8 vertices of a hand group is 3 KB contiguous and
//     becomes 12 packed register pairs per lane;
//   * the host-built skin program (skin_pack) lists, per block, its distinct bones with a dense 8-vector of weights per bone
//     (443 entries synthetic / 386 MANO); every entry updates all 8 vertices DENSE with packed fp32 FFMA2 (48 per entry) —
//     at 59 % density the branches of a skip-if-zero form cost as many issue slots as the zeros;
//   * bone transforms of the hand group are resident in six 1.5 KB shared-memory slots per warp, (re)loaded by bulk copies
//     on a static Belady schedule computed on the host (89 / 65 copies per sweep instead of one per entry);
//   * results are transposed through a 7.4 KB per-warp tile per 16-vertex segment and leave as one contiguous,
//     sector-aligned run per row of verts[B][778][3] (the row pitch is 24 mod 32 bytes: per-row windows with carry slots).
//
// Backward:  dv_posed_v = sum_k w_vk R'_k^T g_v  and
// dA'_k = sum_v w_vk g_v (x) [v_posed_v ; 1].  Two warps share a hand group and a g_verts tile:
// role 0 computes dv_posed (emitted as bf16 hi+mid UMMA tiles for the tcgen05 gradient contraction,
// or fp32 hand-minor), role 1 accumulates the per-bone 3x4 sums in registers per (block, bone) and
// folds them into a 24 KB shared accumulator that only it touches — no atomics.
#include <cuda_bf16.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "common.cuh"
#include "blend_tc.cuh"
#include <stdlib.h>
#include "skin.cuh"
#include "ptx.cuh"

namespace mb {
namespace {

constexpr int SEG_F = SK_SEG * 3;                 // 48 floats per segment row piece
constexpr int TP = 34;                            // tile pitch (floats): element (float f, hand h) at f * 34 + h; conflict-free for
                                                  // lane = hand (compute side) and for the 4 rows x 8 float2 row-piece mapping
constexpr int TILE_FLOATS = SEG_F * TP;           // 1632 floats = 6.4 KB
constexpr int XBLK_FLOATS = SK_BC * 32;           // one block of rest-pose coordinates of a hand group: 3 KB, contiguous
constexpr int PAD_SLOT = SK_SEG - 1;
constexpr int GROUP_BONE_FLOATS = NJ * BONE_F * 32;
constexpr size_t GROUP_V_FLOATS = (size_t)SK_NCOORD * 32;
constexpr int N_TIP = 5;
__constant__ int c_tip_vert[N_TIP] = {333, 444, 672, 555, 745};
__constant__ int c_tip_slot[N_TIP] = {4, 8, 12, 16, 20};

// Streamed data (rest-pose blocks, upstream gradients, vertices, gradient tiles) passes through L2
// with evict-first priority; the bone transforms — re-read ~29 times per sweep, 100 MB for 131 072
// hands — are loaded evict-last.  [round-1 ncu: without the hints the streams pushed the bones out of
// L2 between uses: DRAM reads 31 % above the algorithmic bytes, half of all stalls on bone loads]
struct L2Policies { uint64_t stream, keep; };
__device__ __forceinline__ L2Policies make_policies() { return {l2_policy_evict_first(), l2_policy_evict_last()}; }

__device__ __forceinline__ float ld_stream(const float* p, uint64_t pol) {
    float r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol) : "memory");
    return r;
}
__device__ __forceinline__ float2 ld_stream2(const float* p, uint64_t pol) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(r.x), "=f"(r.y) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float ld_keep(const float* p, uint64_t pol) {
    float r;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_stream(float* p, float v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream2(float* p, const float2& v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" :: "l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream4u(void* p, const uint4& v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" :: "r"(id) : "memory"); }

// shared-memory copy of the skin program (23.7 KB), with the address arithmetic folded in
struct SkinProg {
    int blk_ptr[SK_NBLK + 1];
    int ent_code[SK_MAX_ENT];                    // bone | slot << 4 | wait << 7 (slot schedule of skin_pack)
    int cmd[SK_MAX_CMD + 1];                     // [0] = count, then slot (re)load commands sorted by `after`
    int split[SK_NSEG][8];                       // split sweeps: 6 slot bones, command cursor, bones-touched mask per segment
    alignas(16) float ent_w[SK_MAX_ENT][SK_BV];
    alignas(16) int voff[SK_NPOS];               // float offset of the vertex' x inside a tile: vl * 3 * TP
};
constexpr size_t PROG_BYTES = (sizeof(SkinProg) + 127) & ~size_t(127);

__device__ __forceinline__ void stage_prog(SkinProg& P, const void* blob) {
    const BlobLayout L = blob_layout();
    const int* bp = blob_ptr<int>(blob, L.sk_blk_ptr);
    const int* eb = blob_ptr<int>(blob, L.sk_ent_bone);
    const float* ew = blob_ptr<float>(blob, L.sk_ent_w);
    const uint8_t* vl = blob_ptr<uint8_t>(blob, L.sk_vloc);
    // asynchronous copies (cp.async): the ~6 000 words of the program arrive in one round trip instead of ~24
    // dependent ones per thread — it matters for the small-batch launches, where the sweep itself is 40-80 us
    auto acopy = [](void* dst, const void* src) {
        cp_async4(reinterpret_cast<float*>(dst), reinterpret_cast<const float*>(src));
    };
    for (int i = threadIdx.x; i <= SK_NBLK; i += blockDim.x) acopy(&P.blk_ptr[i], &bp[i]);
    const int ne = bp[SK_NBLK];
    for (int i = threadIdx.x; i < ne; i += blockDim.x) acopy(&P.ent_code[i], &eb[i]);
    const int* cm = blob_ptr<int>(blob, L.sk_cmd);
    const int ncmd = cm[0];
    for (int i = threadIdx.x; i <= ncmd; i += blockDim.x) acopy(&P.cmd[i], &cm[i]);
    const int* sp = blob_ptr<int>(blob, L.sk_split);
    for (int i = threadIdx.x; i < SK_NSEG * 8; i += blockDim.x) acopy(&(&P.split[0][0])[i], &sp[i]);
    for (int i = threadIdx.x; i < ne * SK_BV; i += blockDim.x) acopy(&(&P.ent_w[0][0])[i], &ew[i]);
    // padding positions (only in the last block, whose segment uses 10 of its 16 vertex slots) are
    // parked on the last slot of the tile: written / read like any vertex, never stored, weights all zero
    for (int i = threadIdx.x; i < SK_NPOS; i += blockDim.x) P.voff[i] = (vl[i] == 255 ? PAD_SLOT : vl[i]) * (3 * TP);
    cp_async_wait_all();                                      // the callers' __syncthreads() follows
}

__device__ __forceinline__ void load_w(const SkinProg& P, int e, float (&w)[SK_BV]) {
    const float4 w0 = *reinterpret_cast<const float4*>(&P.ent_w[e][0]);
    const float4 w1 = *reinterpret_cast<const float4*>(&P.ent_w[e][4]);
    w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
}
__device__ __forceinline__ void load_voff(const SkinProg& P, int blk, int (&o)[SK_BV]) {
    const int4 a = *reinterpret_cast<const int4*>(&P.voff[blk * SK_BV]);
    const int4 b = *reinterpret_cast<const int4*>(&P.voff[blk * SK_BV + 4]);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

// ------------------------------------------------------------------ bulk-copy rings
// Everything a sweep reads is contiguous per item and its order is static: a block of rest-pose
// coordinates of a hand group is 24 rows x 128 B = 3 KB of v_posed_t, the bone transform of an entry is
// 12 rows x 128 B = 1.5 KB of bone_t, and the entry list is the program.  Each warp therefore lets the
// bulk-copy (TMA) engine stream both into small shared-memory rings, several items ahead of their
// use and across block and group boundaries, completing on per-slot mbarriers: global-memory and L2
// latency (there is no L1 to speak of — shared memory takes the whole carve-out) never reaches the
// register scoreboard.  [round-1 ncu: with register prefetch one entry ahead half of all stall
// samples were long-scoreboard waits on the first use of a bone or coordinate]
// A wait that cannot hang the GPU: a broken schedule / pipeline traps after ~2 s instead of spinning forever.
__device__ __forceinline__ void mbar_wait_or_trap(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

template <int STAGES, int SLOT_FLOATS>
struct alignas(128) Ring {
    alignas(128) float slot[STAGES][SLOT_FLOATS];
    alignas(8) unsigned long long full[STAGES];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&full[s]), 1);
    }
};
struct Cursor {               // per-warp stream position; every lane keeps an identical copy
    int g, i;                 // next item to request: group g, item i of n
    unsigned issued, consumed;
};
// request the next item of the stream (no-op once the warp's groups are exhausted); src(g, i) is its address
template <int STAGES, int SLOT_FLOATS, class SrcFn>
__device__ __forceinline__ void ring_request(Ring<STAGES, SLOT_FLOATS>& R, Cursor& C, int ngroups, int gstep, int n, int lane,
                                             SrcFn src, uint64_t policy) {
    if (C.g >= ngroups) return;
    if (elect_one()) {
        const unsigned st = C.issued % STAGES;
        const uint32_t bar = smem_u32(&R.full[st]);
        mbar_expect_tx(bar, SLOT_FLOATS * 4);
        bulk_g2s_hint(smem_u32(R.slot[st]), src(C.g, C.i), SLOT_FLOATS * 4, bar, policy);
    }
    ++C.issued;
    if (++C.i == n) { C.i = 0; C.g += gstep; }
}
// wait for the oldest outstanding item; returns its slot
template <int STAGES, int SLOT_FLOATS>
__device__ __forceinline__ const float* ring_wait(Ring<STAGES, SLOT_FLOATS>& R, const Cursor& C) {
    const unsigned st = C.consumed % STAGES;
    const uint32_t bar = smem_u32(&R.full[st]);
    const uint32_t parity = (C.consumed / STAGES) & 1;
    mbar_wait_or_trap(bar, parity);
    return R.slot[st];
}

constexpr int XSTAGES = 2;
constexpr int BONE_SLOT_FLOATS = BONE_F * 32;      // one bone of a hand group: [lane][12], 1.5 KB contiguous in bone_t
typedef Ring<XSTAGES, XBLK_FLOATS> XRing;

// ------------------------------------------------------------------ resident bones
// SK_SLOTS bone transforms of the warp's hand group stay in shared memory; skin_pack's static
// schedule says which slot an entry reads, when that is the first read after a (re)load (wait on the
// slot's mbarrier) and after which entry a slot is refilled (bulk copy issued by an elected lane).
struct alignas(128) BoneCache {
    alignas(128) float slot[SK_SLOTS][BONE_SLOT_FLOATS];
    alignas(8) unsigned long long full[SK_SLOTS];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < SK_SLOTS; ++s) mbar_init(smem_u32(&full[s]), 1);
    }
};
// ci / next_after: cursor into the command list; parity: mbarrier phase per slot; pending: slots whose
// latest load has not been waited for yet (set at issue, cleared by the first read)
struct CacheState { int ci, next_after; unsigned parity, pending; };    // identical in every lane
__device__ __forceinline__ void cache_issue(BoneCache& K, CacheState& S, int c, const float* __restrict__ group_base, uint64_t pol) {
    const int slot = (c >> 10) & 7, bone = (c >> 13) & 15;
    S.pending |= 1u << slot;
    __syncwarp();                                             // every lane is done reading the slot's previous occupant
    if (elect_one()) {
        const uint32_t bar = smem_u32(&K.full[slot]);
        mbar_expect_tx(bar, BONE_SLOT_FLOATS * 4);
        bulk_g2s_hint(smem_u32(K.slot[slot]), group_base + bone * BONE_SLOT_FLOATS, BONE_SLOT_FLOATS * 4, bar, pol);
    }
}
// a warp's first group: nothing was prefetched by a previous sweep
__device__ __forceinline__ void cache_prologue(BoneCache& K, CacheState& S, const SkinProg& P, const float* __restrict__ group_base,
                                               uint64_t pol) {
    const int ncmd = P.cmd[0];
    for (int i = 0; i < ncmd; ++i) {
        const int c = P.cmd[1 + i];
        if ((c >> 17) & 1) cache_issue(K, S, c, group_base, pol);
    }
}
// split sweeps: start at segment `part` of a group from its snapshot — load what the schedule has resident there
__device__ __forceinline__ void cache_begin_part(BoneCache& K, CacheState& S, const SkinProg& P, int part,
                                                 const float* __restrict__ group_base, uint64_t pol) {
#pragma unroll
    for (int sl = 0; sl < SK_SLOTS; ++sl) {
        const int b = P.split[part][sl];
        if (b >= 0) cache_issue(K, S, (sl << 10) | (b << 13), group_base, pol);
    }
    S.ci = P.split[part][6];
    S.next_after = S.ci < P.cmd[0] ? (P.cmd[1 + S.ci] & 1023) : (1 << 30);
}
// split sweeps: nothing may stay in flight when the warp moves on to another unit
__device__ __forceinline__ void cache_drain(BoneCache& K, CacheState& S) {
#pragma unroll
    for (int sl = 0; sl < SK_SLOTS; ++sl)
        if (S.pending & (1u << sl)) {
            mbar_wait_or_trap(smem_u32(&K.full[sl]), (S.parity >> sl) & 1);
            S.parity ^= 1u << sl;
        }
    S.pending = 0;
}
__device__ __forceinline__ void cache_begin_group(CacheState& S, const SkinProg& P) {
    S.ci = 0;
    S.next_after = P.cmd[0] > 0 ? (P.cmd[1] & 1023) : (1 << 30);
}
// after entry e has been consumed: issue the loads scheduled behind it (group_base = this group's
// bones; the next group's are next_off floats further when it exists)
__device__ __forceinline__ void cache_after_entry(BoneCache& K, CacheState& S, const SkinProg& P, int e,
                                                  const float* __restrict__ group_base, bool has_next, size_t next_off,
                                                  uint64_t pol) {
    while (S.next_after == e + 1) {
        const int c = P.cmd[1 + S.ci];
        if ((c >> 17) & 1) { if (has_next) cache_issue(K, S, c, group_base + next_off, pol); }
        else cache_issue(K, S, c, group_base, pol);
        ++S.ci;
        S.next_after = S.ci < P.cmd[0] ? (P.cmd[1 + S.ci] & 1023) : (1 << 30);
    }
}
// the slot of an entry, ready to read (lane = hand: 48 bytes at lane * 12)
__device__ __forceinline__ const float4* cache_entry(BoneCache& K, CacheState& S, int code, int lane) {
    const int slot = (code >> 4) & 7;
    if (S.pending & (1u << slot)) {                           // first read since the slot's last (re)load
        const uint32_t bar = smem_u32(&K.full[slot]);
        const uint32_t par = (S.parity >> slot) & 1;
        mbar_wait_or_trap(bar, par);
        S.parity ^= 1u << slot;
        S.pending &= ~(1u << slot);
    }
    return reinterpret_cast<const float4*>(K.slot[slot] + lane * BONE_F);
}

// ------------------------------------------------------------------ forward
// Packed fp32 (FFMA2, sm_100): two vertices of the block per instruction, the bone element is the
// broadcast scalar operand.  X[c][m] = coordinate c of vertices (2m, 2m+1).  Every entry is
// computed DENSE over the block's 8 vertices (zero weights contribute exact zeros): with the packed
// instruction that is 48 FFMA2 per entry, fewer issue slots than the skip-if-zero variant spent on
// tests, branches and reconvergence barriers alone, and it is straight-line code.
__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }
__device__ __forceinline__ void fma_entry2(const float (&A)[BONE_F], const float (&w)[SK_BV],
                                           const float2 (&X)[3][4], float2 (&ACC)[3][4]) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const float2 wp = make_float2(w[2 * m], w[2 * m + 1]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float2 t = __ffma2_rn(X[2][m], bc(A[4 * c + 2]), bc(A[4 * c + 3]));
            t = __ffma2_rn(X[1][m], bc(A[4 * c + 1]), t);
            t = __ffma2_rn(X[0][m], bc(A[4 * c]), t);
            ACC[c][m] = __ffma2_rn(wp, t, ACC[c][m]);
        }
    }
}
constexpr int SKF_WARPS = 8;                       // autonomous warps per CTA; 1 CTA per SM (the kernel is bound by the
                                                   // memory system, not by warp count: 7..20 warps measured the same)
constexpr int CARRY_F = 8;                         // floats of the previous segment kept in front of the tile
constexpr int SKF_THREADS = SKF_WARPS * 32;
struct alignas(128) FwdWarpShared {
    BoneCache bones;
    XRing xs;
    alignas(16) float tile[(CARRY_F + SEG_F) * TP];          // slots 0..7: tail of the previous segment, 8..55: this segment
};
constexpr size_t SKF_SMEM = PROG_BYTES + (size_t)SKF_WARPS * sizeof(FwdWarpShared);

// Row pieces of a 16-vertex segment <-> tile: one warp instruction covers 4 rows x 8 float2 (64 B per row).
// lane -> (r = lane >> 3, p = lane & 7); piece block qb (0..2), row block rb (0..7):
//   float f = 16 qb + 2 p + e, hand h = 4 rb + r   ->  tile[f * TP + h]
struct RowMap {
    int r, p;
    __device__ __forceinline__ RowMap(int lane) : r(lane >> 3), p(lane & 7) {}
    __device__ __forceinline__ int tile_base() const { return (2 * p) * TP + r; }
    __device__ __forceinline__ size_t row_base() const { return (size_t)r * NVC + 2 * p; }
};

// block of 8 vertices as packed pairs: X[c][m] = (x_{2m,c}, x_{2m+1,c}), straight from global memory ...
__device__ __forceinline__ void load_xpairs(float2 (&X)[3][4], const float* __restrict__ vb, uint64_t pol) {
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            X[c][m] = make_float2(ld_stream(vb + (6 * m + c) * 32, pol), ld_stream(vb + (6 * m + 3 + c) * 32, pol));
}
// ... or out of a ring slot
__device__ __forceinline__ void slot_to_pairs(float2 (&X)[3][4], const float* sl) {
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int c = 0; c < 3; ++c) X[c][m] = make_float2(sl[(6 * m + c) * 32], sl[(6 * m + 3 + c) * 32]);
}

// a finished 16-vertex segment leaves the tile as 192-byte row pieces (kept out of line: once per segment)
// A finished segment leaves the tile as one contiguous run per row (24 lanes x 8 B; four rows take three
// instructions, flat index i = lane + 32 it -> (row i / 24, float2 i % 24)), and every run is SECTOR-ALIGNED:
// rows of verts[B][778][3] are 9 336 B = 24 mod 32 apart, so row h's 32-byte boundaries sit at floats
// f = 2h mod 8.  Row h therefore stores the window [48 seg - d, 48 seg + 48 - d), d = (0, 6, 4, 2)[h % 4] floats
// — the last d floats of the previous segment (kept in the tile's carry slots) instead of the last d of this
// one — except at the two ends of the row.  [measured: a 32-byte-aligned row pitch alone was worth 12 %:
// partial-sector writes cost L2 a read-modify-write and DRAM 1.1 GB of fill reads per 2^20 hands]
// A split sweep (part_first / part_last) has no carry at its first segment — the window starts at the
// segment — and writes the d floats the next part will not at its last one.
template <bool SPLIT>
__device__ __noinline__ void store_segment(const float* tile, float* row0 /* verts row 0 of the group */, int seg,
                                           int nh, int lane, uint64_t pol, bool part_first = false, bool part_last = false) {
    const bool last = seg == SK_NSEG - 1;
    if (SPLIT && part_last && !last && lane < nh) {
        const int d2 = (4 - lane) & 3;
        for (int q = 0; q < d2; ++q) {
            const float* t = tile + (CARRY_F + SEG_F - 2 * d2 + 2 * q) * TP + lane;
            st_stream2(row0 + (size_t)lane * NVC + SEG_F * seg + SEG_F - 2 * d2 + 2 * q, make_float2(t[0], t[TP]), pol);
        }
    }
#pragma unroll 1
    for (int rb = 0; rb < 8; ++rb) {
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int i = lane + 32 * it;
            const int rr = i / 24, pp = i - rr * 24;
            const int h = rb * 4 + rr;
            const int d2 = (4 - rr) & 3;                       // d / 2 = (0, 3, 2, 1)[h % 4]; h % 4 == rr
            const int f2 = 24 * seg - d2 + pp;                 // float2 index inside the row
            const int np = last ? (NV * 3 / 2 - 24 * seg + d2) : 24;
            if (h < nh && pp < np && f2 >= (SPLIT && part_first ? 24 * seg : 0)) {
                const float* t = tile + (CARRY_F - 2 * d2 + 2 * pp) * TP + h;
                st_stream2(row0 + (size_t)h * NVC + 2 * f2, make_float2(t[0], t[TP]), pol);
            }
        }
    }
}

// SPLIT: a work unit is `spu` segments of a group's sweep (spu divides 49) instead of the whole sweep, so a
// batch of fewer groups than resident warps still fills the machine [B = 4096: 158 us -> see ncu_history];
// the resident bones of a part come from skin_pack's snapshot of the slot schedule at its first segment.
template <bool SPLIT>
__global__ void __launch_bounds__(SKF_THREADS, 1)
skin_forward_kernel(const void* __restrict__ blob, const float* __restrict__ v_posed_t,
                    const float* __restrict__ bone_t, int B, float* __restrict__ verts, float* __restrict__ joints, int spu) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SkinProg& P = *reinterpret_cast<SkinProg*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FwdWarpShared& W = reinterpret_cast<FwdWarpShared*>(smem_raw + PROG_BYTES)[warp];
    if (lane == 0) {
        W.bones.init();
        W.xs.init();
        fence_barrier_init();
    }
    stage_prog(P, blob);
    __syncthreads();
    float* tl = W.tile + CARRY_F * TP + lane;                 // compute side: (float f of the segment, lane)
    const int PARTS = SPLIT ? skin_units_per_group(spu) : 1;  // work units per group (the last one may be short)
    const int UBLK = SPLIT ? spu * SK_SEG_BLKS : SK_NBLK;     // blocks per full unit
    const int nunits = ((B + 31) >> 5) * PARTS;
    const int u0 = blockIdx.x + warp * gridDim.x, ustep = gridDim.x * SKF_WARPS;
    auto x_src = [&](int u, int i) {
        return v_posed_t + (size_t)(u / PARTS) * GROUP_V_FLOATS + (size_t)((u % PARTS) * UBLK + i) * XBLK_FLOATS;
    };
    auto unit_blocks = [&](int u) {                           // of the unit the ring's request cursor is in
        if (!SPLIT) return SK_NBLK;
        const int b0 = (u % PARTS) * UBLK;
        return (b0 + UBLK < SK_NBLK ? b0 + UBLK : SK_NBLK) - b0;
    };
    const L2Policies pol = make_policies();
    Cursor CX = {u0, 0, 0u, 0u};
#pragma unroll
    for (int s = 0; s < XSTAGES; ++s) ring_request(W.xs, CX, nunits, ustep, unit_blocks(CX.g), lane, x_src, pol.stream);
    CacheState CS = {0, 0, 0u, 0u};
    if (!SPLIT && u0 < nunits) cache_prologue(W.bones, CS, P, bone_t + (size_t)u0 * GROUP_BONE_FLOATS, pol.keep);
    const size_t next_off = (size_t)ustep * GROUP_BONE_FLOATS;

    for (int u = u0; u < nunits; u += ustep) {
        const int g = u / PARTS, part = u % PARTS;
        const int nh = (B - g * 32) < 32 ? (B - g * 32) : 32;
        const float* bgrp = bone_t + (size_t)g * GROUP_BONE_FLOATS;
        const bool has_next = !SPLIT && u + ustep < nunits;
        if (SPLIT) cache_begin_part(W.bones, CS, P, part * spu, bgrp, pol.keep);
        else cache_begin_group(CS, P);
        const int blk1 = (part + 1) * UBLK < SK_NBLK ? (part + 1) * UBLK : SK_NBLK;
#pragma unroll 1
        for (int blk = part * UBLK; blk < blk1; ++blk) {
            float2 X[3][4], ACC[3][4];
            slot_to_pairs(X, ring_wait(W.xs, CX) + lane);
            ++CX.consumed;
            __syncwarp();                                     // every lane has its copy: the slot can be refilled
            ring_request(W.xs, CX, nunits, ustep, unit_blocks(CX.g), lane, x_src, pol.stream);
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int m = 0; m < 4; ++m) ACC[c][m] = make_float2(0.f, 0.f);
            const int e1 = P.blk_ptr[blk + 1];
#pragma unroll 1
            for (int e = P.blk_ptr[blk]; e < e1; ++e) {
                float w[SK_BV];
                load_w(P, e, w);
                const float4* sl = cache_entry(W.bones, CS, P.ent_code[e], lane);
                const float4 a0 = sl[0], a1 = sl[1], a2 = sl[2];
                cache_after_entry(W.bones, CS, P, e, bgrp, has_next, next_off, pol.keep);   // may refill the slot just read
                const float A[BONE_F] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
                fma_entry2(A, w, X, ACC);
            }
            int vo[SK_BV];
            load_voff(P, blk, vo);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                float* t0 = tl + vo[2 * m];
                float* t1 = tl + vo[2 * m + 1];
                t0[0] = ACC[0][m].x; t0[TP] = ACC[1][m].x; t0[2 * TP] = ACC[2][m].x;
                t1[0] = ACC[0][m].y; t1[TP] = ACC[1][m].y; t1[2 * TP] = ACC[2][m].y;
            }
            if (blk & 1) {                                    // second block of a segment: the segment is complete
                const int seg = blk >> 1;
                __syncwarp();
                if (SPLIT)
                    store_segment<true>(W.tile, verts + (size_t)g * 32 * NVC, seg, nh, lane, pol.stream,
                                  blk == part * UBLK + 1, blk == blk1 - 1);
                else
                    store_segment<false>(W.tile, verts + (size_t)g * 32 * NVC, seg, nh, lane, pol.stream);
                if (joints != nullptr && lane < nh) {
#pragma unroll
                    for (int t = 0; t < N_TIP; ++t)
                        if (c_tip_vert[t] / SK_SEG == seg) {
                            const float* tv = tl + (c_tip_vert[t] % SK_SEG) * (3 * TP);
                            float* o = joints + ((size_t)g * 32 + lane) * (NOUTJ * 3) + c_tip_slot[t] * 3;
                            o[0] = tv[0]; o[1] = tv[TP]; o[2] = tv[2 * TP];
                        }
                }
                __syncwarp();
                // keep the segment's last 8 floats of this lane's hand for the next segment's shifted windows
#pragma unroll
                for (int c = 0; c < CARRY_F; ++c) tl[(c - CARRY_F) * TP] = tl[(SEG_F - CARRY_F + c) * TP];
                __syncwarp();
            }
        }
        if (SPLIT) cache_drain(W.bones, CS);
    }
}

// ----------------------------------------------------------------- backward
constexpr int SKB_PAIRS = 4;
constexpr int SKB_THREADS = SKB_PAIRS * 64;
constexpr int DP = 33;                             // accumulator pitch: element (e, hand) at e * 33 + hand
struct alignas(128) BwdPairShared {
    BoneCache bones;                                         // role 0
    alignas(16) float tile[TILE_FLOATS];                     // upstream gradient of one 16-vertex segment, transposed
    alignas(16) float dacc[NJ * BONE_F * DP];                // per-bone 3x4 sums of the group (role 1 only)
};
constexpr size_t SKB_SMEM = PROG_BYTES + (size_t)SKB_PAIRS * sizeof(BwdPairShared);

// gather the upstream gradient of a block's 8 vertices from the segment tile as packed pairs:
// G[c][m] = (g_{2m,c}, g_{2m+1,c})
__device__ __forceinline__ void gather_block(const SkinProg& P, const float* tl, int blk, float2 (&G)[3][4]) {
    int vo[SK_BV];
    load_voff(P, blk, vo);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const float* t0 = tl + vo[2 * m];
        const float* t1 = tl + vo[2 * m + 1];
#pragma unroll
        for (int c = 0; c < 3; ++c) G[c][m] = make_float2(t0[c * TP], t1[c * TP]);
    }
}

// role 0: DV[c][m] += w (R^T g)_c for the vertex pairs of the block, dense (FFMA2); A holds the first
// 11 elements of the 3x4 transform (R = A[0..2], A[4..6], A[8..10])
__device__ __forceinline__ void dv_entry2(const float (&A)[11], const float (&w)[SK_BV],
                                          const float2 (&G)[3][4], float2 (&DV)[3][4]) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const float2 wp = make_float2(w[2 * m], w[2 * m + 1]);
        const float2 wx = __fmul2_rn(wp, G[0][m]), wy = __fmul2_rn(wp, G[1][m]), wz = __fmul2_rn(wp, G[2][m]);
#pragma unroll
        for (int c = 0; c < 3; ++c)
            DV[c][m] = __ffma2_rn(wx, bc(A[c]), __ffma2_rn(wy, bc(A[4 + c]), __ffma2_rn(wz, bc(A[8 + c]), DV[c][m])));
    }
}

// role 1: dacc[k] += sum_j w_jk g_j (x) [v_j ; 1] for every bone k of the block (dl = dacc + lane);
// even / odd vertices accumulate in the two halves of packed registers and are added at the end
__device__ __forceinline__ void skin_block_da(const SkinProg& P, int blk, float* dl,
                                              const float2 (&G)[3][4], const float2 (&V)[3][4]) {
    const int e1 = P.blk_ptr[blk + 1];
#pragma unroll 1
    for (int e = P.blk_ptr[blk]; e < e1; ++e) {
        float w[SK_BV];
        load_w(P, e, w);
        float* d = dl + (P.ent_code[e] & 15) * (BONE_F * DP);
        float2 a[BONE_F];
#pragma unroll
        for (int i = 0; i < BONE_F; ++i) a[i] = make_float2(d[i * DP], 0.f);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const float2 wp = make_float2(w[2 * m], w[2 * m + 1]);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float2 wg = __fmul2_rn(wp, G[r][m]);
                a[4 * r]     = __ffma2_rn(wg, V[0][m], a[4 * r]);
                a[4 * r + 1] = __ffma2_rn(wg, V[1][m], a[4 * r + 1]);
                a[4 * r + 2] = __ffma2_rn(wg, V[2][m], a[4 * r + 2]);
                a[4 * r + 3] = __fadd2_rn(a[4 * r + 3], wg);
            }
        }
#pragma unroll
        for (int i = 0; i < BONE_F; ++i) d[i * DP] = a[i].x + a[i].y;
    }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&p);
}

// SPLIT: a work unit is `spu` segments of a group's sweep; the unit's bone sums go to dparts[unit][192][32]
// (touched bones only) and dbone_reduce_kernel adds the units of a group in a fixed order.
template <bool SPLIT>
__global__ void __launch_bounds__(SKB_THREADS, 1)
skin_backward_kernel(const void* __restrict__ blob, const float* __restrict__ v_posed_t,
                     const float* __restrict__ bone_t, const float* __restrict__ g_verts,
                     const float* __restrict__ g_joints, int B,
                     float* __restrict__ dv_t, unsigned char* __restrict__ dvp, float* __restrict__ dbone,
                     int dbone_hand_minor, float* __restrict__ dparts, int spu, int role_flip) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SkinProg& P = *reinterpret_cast<SkinProg*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // warps w and w + 4 share a scheduler: with role = warp & 1 two schedulers would run only role-0 warps and two only the
    // (longer) role-1 warps; flipping the roles of the upper four warps gives every scheduler one of each
    const int pair = warp >> 1, role = (warp & 1) ^ (role_flip & (warp >> 2) & 1);
    // timing experiments (MANO_B200_SKB_EXP, profiles/tools/skb_exp.sh; the results are WRONG with any bit set): 1 / 2 = role 0 /
    // role 1 skips its arithmetic, 4 = no gradient-tile stores, 8 / 16 = role 1 / role 0 only takes part in the tile hand-over
    const int exp = role_flip >> 8;
    BwdPairShared& W = reinterpret_cast<BwdPairShared*>(smem_raw + PROG_BYTES)[pair];
    if (role == 0 && lane == 0) {
        W.bones.init();
        fence_barrier_init();
    }
    stage_prog(P, blob);
    __syncthreads();
    float* tl = W.tile + lane;
    const RowMap rm(lane);
    float* tsd = W.tile + rm.tile_base();                     // load side of the g tile (same mapping as the forward's store side)
    const int bar = 1 + pair;
    const int PARTS = SPLIT ? skin_units_per_group(spu) : 1;  // work units per group (the last one may be short)
    const int nunits = ((B + 31) >> 5) * PARTS;
    const int u0 = blockIdx.x + pair * gridDim.x, ustep = gridDim.x * SKB_PAIRS;
    const L2Policies pol = make_policies();
    CacheState CS = {0, 0, 0u, 0u};
    if (!SPLIT && role == 0 && u0 < nunits) cache_prologue(W.bones, CS, P, bone_t + (size_t)u0 * GROUP_BONE_FLOATS, pol.keep);
    const size_t next_off = (size_t)ustep * GROUP_BONE_FLOATS;

    for (int u = u0; u < nunits; u += ustep) {
        const int g = u / PARTS;
        const int seg0 = SPLIT ? (u % PARTS) * spu : 0, seg1 = SPLIT && seg0 + spu < SK_NSEG ? seg0 + spu : SK_NSEG;
        const int nh = (B - g * 32) < 32 ? (B - g * 32) : 32;
        const float* vb = v_posed_t + (size_t)g * GROUP_V_FLOATS + lane;
        const float* bgrp = bone_t + (size_t)g * GROUP_BONE_FLOATS;
        const bool has_next = !SPLIT && u + ustep < nunits;
        if (role == 0) {
            if (SPLIT) cache_begin_part(W.bones, CS, P, seg0, bgrp, pol.keep);
            else cache_begin_group(CS, P);
        }
        const float* grow = g_verts + (size_t)g * 32 * NVC + rm.row_base();
        // each role loads half of a segment's row pieces (one instruction = 4 rows x 8 float2) into
        // registers one segment ahead of its use
        float2 pre[4][3];
        auto prefetch_g = [&](int seg) {
            const int nf = (seg == SK_NSEG - 1) ? (NV - seg * SK_SEG) * 3 : SEG_F;
            const float* src = grow + seg * SEG_F;
#pragma unroll
            for (int rb2 = 0; rb2 < 4; ++rb2) {
                const int rb = rb2 * 2 + role;
#pragma unroll
                for (int qb = 0; qb < 3; ++qb) {
                    const bool ok = (rb * 4 + rm.r < nh) && (qb * 16 + 2 * rm.p < nf);
                    pre[rb2][qb] = ok ? ld_stream2(src + (size_t)rb * 4 * NVC + qb * 16, pol.stream) : make_float2(0.f, 0.f);
                }
            }
        };
        prefetch_g(seg0);
        float2 va[3][4], vbk[3][4];                           // role 1: rest-pose blocks (packed pairs), loaded one block ahead
        if (role == 0) {
            if (dvp != nullptr && seg1 == SK_NSEG) {
                // zero the 16 padding K columns (2352..2367) of the last chunk of the gradient tiles
                unsigned char* tb = dvp + (size_t)(g >> 2) * TCB_A_TILE_BYTES + (size_t)(TCB_K_CHUNKS - 1) * TCB_A_CHUNK_BYTES;
                const int rg = (g & 3) * 4 + (lane >> 3), r = lane & 7;
#pragma unroll
                for (int kg = 2; kg < 4; ++kg) {
                    unsigned char* d = tb + ((rg * 4 + kg) * 8 + r) * 16;
                    st_stream4u(d, make_uint4(0, 0, 0, 0), pol.stream);
                    st_stream4u(d + TC_A_BLOCK_BYTES, make_uint4(0, 0, 0, 0), pol.stream);
                }
            }
        } else {
            load_xpairs(va, vb + (size_t)(2 * seg0) * XBLK_FLOATS, pol.stream);
#pragma unroll 8
            for (int i = 0; i < NJ * BONE_F; ++i) W.dacc[i * DP + lane] = 0.f;
        }

        // role 0, one block: dv_posed of its 8 vertices -> gradient tiles (or fp32 hand-minor)
        auto dv_block = [&](int blk) {
            float2 G[3][4], DV[3][4];
            gather_block(P, tl, blk, G);
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int m = 0; m < 4; ++m) DV[c][m] = make_float2(0.f, 0.f);
            const int e1 = P.blk_ptr[blk + 1];
#pragma unroll 1
            for (int e = P.blk_ptr[blk]; e < e1; ++e) {
                float A[11], w[SK_BV];
                load_w(P, e, w);
                {
                    const float4* sl = cache_entry(W.bones, CS, P.ent_code[e], lane);
                    const float4 a0 = sl[0], a1 = sl[1], a2 = sl[2];
                    A[0] = a0.x; A[1] = a0.y; A[2] = a0.z; A[3] = a0.w; A[4] = a1.x; A[5] = a1.y; A[6] = a1.z; A[7] = a1.w;
                    A[8] = a2.x; A[9] = a2.y; A[10] = a2.z;
                    cache_after_entry(W.bones, CS, P, e, bgrp, has_next, next_off, pol.keep);
                }
                if (!(exp & 1)) dv_entry2(A, w, G, DV);
            }
            float dv[SK_BC];                                    // block order: dv[3j + c]
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int c = 0; c < 3; ++c) { dv[6 * m + c] = DV[c][m].x; dv[6 * m + 3 + c] = DV[c][m].y; }
            if (exp & 4) return;
            if (dvp != nullptr) {
                // A operand of the tcgen05 gradient contraction: bf16 hi + mid, UMMA canonical K-major
                // blocks; a lane owns a tile row, so 8 consecutive K values are one 16-byte group and
                // 8 lanes write one contiguous 128-byte core matrix.
                unsigned char* tb = dvp + (size_t)(g >> 2) * TCB_A_TILE_BYTES;
                const int rg = (g & 3) * 4 + (lane >> 3), r = lane & 7;
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    uint32_t hi[4], mid[4];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float a = dv[t * 8 + 2 * p], b = dv[t * 8 + 2 * p + 1];
                        hi[p] = pack_bf16x2(a, b);
                        const float ah = __uint_as_float(hi[p] << 16), bh = __uint_as_float(hi[p] & 0xffff0000u);
                        mid[p] = pack_bf16x2(a - ah, b - bh);
                    }
                    const int kg8 = blk * 3 + t;
                    unsigned char* d = tb + (size_t)(kg8 >> 2) * TCB_A_CHUNK_BYTES + ((rg * 4 + (kg8 & 3)) * 8 + r) * 16;
                    st_stream4u(d, make_uint4(hi[0], hi[1], hi[2], hi[3]), pol.stream);
                    st_stream4u(d + TC_A_BLOCK_BYTES, make_uint4(mid[0], mid[1], mid[2], mid[3]), pol.stream);
                }
            } else {
                float* dg = dv_t + (size_t)g * GROUP_V_FLOATS + (size_t)blk * XBLK_FLOATS + lane;
#pragma unroll
                for (int i = 0; i < SK_BC; ++i) st_stream(dg + i * 32, dv[i], pol.stream);
            }
        };
        // role 1, one block: per-bone sums
        auto da_block = [&](int blk, const float2 (&V)[3][4]) {
            float2 G[3][4];
            gather_block(P, tl, blk, G);
            if (!(exp & 2)) skin_block_da(P, blk, W.dacc + lane, G, V);
            else if (G[0][0].x + V[0][0].x == 123.456f) W.dacc[lane] = 1.f;
        };

        for (int seg = seg0; seg < seg1; ++seg) {
            pair_barrier(bar);                                  // both warps are done with the previous tile
#pragma unroll
            for (int rb2 = 0; rb2 < 4; ++rb2) {
                const int rb = rb2 * 2 + role;
#pragma unroll
                for (int qb = 0; qb < 3; ++qb) {
                    tsd[(qb * 16) * TP + rb * 4] = pre[rb2][qb].x;
                    tsd[(qb * 16 + 1) * TP + rb * 4] = pre[rb2][qb].y;
                }
            }
            pair_barrier(bar);                                  // tile complete
            bool has_tip = false;
#pragma unroll
            for (int t = 0; t < N_TIP; ++t) has_tip |= (c_tip_vert[t] / SK_SEG == seg);
            if (has_tip) {
                // fingertip joints are vertices: their upstream gradient joins g_verts (A.2 step 1)
                if (role == 0 && lane < nh) {
#pragma unroll
                    for (int t = 0; t < N_TIP; ++t)
                        if (c_tip_vert[t] / SK_SEG == seg) {
                            float* tv = tl + (c_tip_vert[t] % SK_SEG) * (3 * TP);
                            const float* gj = g_joints + ((size_t)g * 32 + lane) * (NOUTJ * 3) + c_tip_slot[t] * 3;
                            tv[0] += gj[0]; tv[TP] += gj[1]; tv[2 * TP] += gj[2];
                        }
                }
                pair_barrier(bar);
            }
            if (seg + 1 < seg1) prefetch_g(seg + 1);
            if (role == 0) {
                if (!(exp & 16)) {
                    dv_block(2 * seg);
                    dv_block(2 * seg + 1);
                }
            } else if (exp & 8) {
            } else {
                load_xpairs(vbk, vb + (size_t)(2 * seg + 1) * XBLK_FLOATS, pol.stream);
                da_block(2 * seg, va);
                if (seg + 1 < seg1) load_xpairs(va, vb + (size_t)(2 * seg + 2) * XBLK_FLOATS, pol.stream);
                da_block(2 * seg + 1, vbk);
            }
        }
        if (SPLIT) {
            if (role == 0) {
                cache_drain(W.bones, CS);
            } else {
                __syncwarp();
                int touched = 0;
                for (int seg = seg0; seg < seg1; ++seg) touched |= P.split[seg][7];
                float* dt = dparts + (size_t)u * GROUP_BONE_FLOATS + lane;
                for (int k = 0; k < NJ; ++k)
                    if ((touched >> k) & 1) {
#pragma unroll
                        for (int i = 0; i < BONE_F; ++i) dt[(k * BONE_F + i) * 32] = W.dacc[(k * BONE_F + i) * DP + lane];
                    }
                __syncwarp();
            }
        } else if (role == 1) {
            // per-bone sums of the group leave as dbone[h][16][12] rows (transposed out of the accumulator)
            __syncwarp();
            if (dbone_hand_minor) {                            // the lane = hand pose backward reads them as they are
                float* dt = dbone + (size_t)g * GROUP_BONE_FLOATS + lane;
#pragma unroll 8
                for (int i = 0; i < NJ * BONE_F; ++i) dt[i * 32] = W.dacc[i * DP + lane];
            } else {
                float* drow = dbone + (size_t)g * 32 * (NJ * BONE_F) + lane;
                const float* da = W.dacc + lane * DP;
                for (int h = 0; h < nh; ++h)
#pragma unroll
                    for (int i = 0; i < NJ * BONE_F / 32; ++i) drow[(size_t)h * (NJ * BONE_F) + 32 * i] = da[(32 * i) * DP + h];
            }
            __syncwarp();
        }
    }
}

// split backward sweeps: dbone[g] = sum over the group's units (ascending, touched bones only) of dparts.
// One CTA per (group, bone): thread -> (element i = t / 32, hand = t % 32).
__global__ void __launch_bounds__(BONE_F * 32)
dbone_reduce_kernel(const void* __restrict__ blob, const float* __restrict__ dparts, int B, int spu,
                    float* __restrict__ dbone, int dbone_hand_minor) {
    __shared__ int touched[SK_NSEG], plist[SK_NSEG], np_s;
    const int* split = blob_ptr<int>(blob, blob_layout().sk_split);
    const int g = blockIdx.x / NJ, k = blockIdx.x % NJ;
    const int i = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int parts = skin_units_per_group(spu);
    if (threadIdx.x < SK_NSEG) touched[threadIdx.x] = split[threadIdx.x * 8 + 7];
    __syncthreads();
    if (threadIdx.x == 0) {                                    // the units that hold a sum for bone k, ascending
        int n = 0;
        for (int p = 0; p < parts; ++p) {
            int m = 0;
            for (int seg = p * spu; seg < (p + 1) * spu && seg < SK_NSEG; ++seg) m |= touched[seg];
            if ((m >> k) & 1) plist[n++] = p;
        }
        np_s = n;
    }
    __syncthreads();
    const int n = np_s;
    const float* src = dparts + (size_t)g * parts * GROUP_BONE_FLOATS + (k * BONE_F + i) * 32 + lane;
    float acc = 0.f;
    for (int j = 0; j < n; j += 4) {                           // four independent loads in flight, added in order
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = j + q < n ? src[(size_t)plist[j + q] * GROUP_BONE_FLOATS] : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc += v[q];
    }
    if (dbone_hand_minor) dbone[(size_t)g * GROUP_BONE_FLOATS + (k * BONE_F + i) * 32 + lane] = acc;
    else if (g * 32 + lane < B) dbone[((size_t)g * 32 + lane) * (NJ * BONE_F) + k * BONE_F + i] = acc;
}

// ------------------------------------------------- layout conversions (fp32 anchor mode, mb_lbs_forward)
// rows[B][pitch] (original vertex order) -> t[group][SK_NCOORD][32] (block order, hand-minor)
__global__ void rows_to_t_kernel(const void* __restrict__ blob, const float* __restrict__ rows, int pitch, int B,
                                 float* __restrict__ t) {
    const int* perm = blob_ptr<int>(blob, blob_layout().sk_perm);
    const long long n = (long long)((B + 31) >> 5) * SK_NCOORD * 32;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int lane = (int)(i & 31);
        const long long gc = i >> 5;
        const int c = (int)(gc % SK_NCOORD);
        const long long h = (gc / SK_NCOORD) * 32 + lane;
        const int v = perm[c / 3];
        t[i] = (v >= 0 && h < B) ? rows[h * pitch + v * 3 + (c % 3)] : 0.f;
    }
}
// inverse: t -> rows[B][pitch]; columns that are no vertex coordinate are zeroed
__global__ void t_to_rows_kernel(const void* __restrict__ blob, const float* __restrict__ t, int pitch, int B,
                                 float* __restrict__ rows) {
    const int* perm = blob_ptr<int>(blob, blob_layout().sk_perm);
    const long long n = (long long)((B + 31) >> 5) * SK_NCOORD * 32;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int lane = (int)(i & 31);
        const long long gc = i >> 5;
        const int c = (int)(gc % SK_NCOORD);
        const long long h = (gc / SK_NCOORD) * 32 + lane;
        const int v = perm[c / 3];
        if (v >= 0 && h < B) rows[h * pitch + v * 3 + (c % 3)] = t[i];
        if (c < pitch - NVC && h < B) rows[h * pitch + NVC + c] = 0.f;
    }
}
// bone[B][16][12] -> bone_t[group][16][32][12]
__global__ void bone_rows_to_t_kernel(const float* __restrict__ bone, int B, float* __restrict__ bone_t) {
    const long long n = (long long)((B + 31) >> 5) * GROUP_BONE_FLOATS;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i % BONE_F);                       // bone_t[group][bone][lane][12]
        const int lane = (int)((i / BONE_F) & 31);
        const int k = (int)((i / (BONE_F * 32)) % NJ);
        const long long h = (i / GROUP_BONE_FLOATS) * 32 + lane;
        bone_t[i] = h < B ? bone[h * (NJ * BONE_F) + k * BONE_F + e] : 0.f;
    }
}

inline int conv_grid(long long n) {
    long long b = (n + 255) / 256;
    const long long cap = (long long)NUM_SMS * 16;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

// ---------------------------------------------------------------- host: skin program
// 1. Blocks: every 16-vertex segment is split into two blocks of 8 so that the two blocks' bone sets
//    are as small as possible (exhaustive over the C(16,8)/2 splits: 6 435 per segment).
// 2. Slot schedule: a warp keeps SK_SLOTS bone transforms of its hand group resident in shared
//    memory.  The order in which a sweep touches bones is static, so the replacement policy is
//    Belady's (evict the bone whose next use is farthest) computed here, and every (re)load is issued
//    right after the LAST use of the bone it replaces — typically dozens of entries before its own
//    first use, so the bulk copy's latency is hidden without a deep ring, and a bone is fetched
//    ~5 times per sweep instead of once per (block, bone) entry (~29 times).  [round-1 ncu: the
//    per-entry ring ran the skinning kernels at the L2 bandwidth limit — 58 KB of L2 sector
//    traffic per hand, 21.6 KB of it bone re-reads]
//    Per entry: bone | slot << 4 | wait << 7 (wait = first use after a load: wait on the slot's mbarrier).
//    Commands, sorted by `after`: issue "slot <- bone" once entry `after` has been consumed;
//    next_group marks the loads of the following group's first occupants, issued during this
//    group's tail (and by the prologue for a warp's first group).
int skin_pack(const float* skin_w, const int32_t* skin_b, void* host_blob, int32_t* coord_map) {
    const BlobLayout L = blob_layout();
    char* out = reinterpret_cast<char*>(host_blob);
    int* blk_ptr = reinterpret_cast<int*>(out + L.sk_blk_ptr);
    int* ent_bone = reinterpret_cast<int*>(out + L.sk_ent_bone);
    float* ent_w = reinterpret_cast<float*>(out + L.sk_ent_w);
    uint8_t* vloc = reinterpret_cast<uint8_t*>(out + L.sk_vloc);
    int* perm = reinterpret_cast<int*>(out + L.sk_perm);
    int* cmd = reinterpret_cast<int*>(out + L.sk_cmd);

    std::vector<unsigned> mask(NV, 0u);
    std::vector<float> dense((size_t)NV * NJ, 0.f);
    for (int v = 0; v < NV; ++v)
        for (int s = 0; s < MAX_INFL; ++s) {
            const float w = skin_w[v * MAX_INFL + s];
            const int b = skin_b[v * MAX_INFL + s];
            if (w == 0.f) continue;
            if (b < 0 || b >= NJ) return MB_E_MODEL;
            mask[v] |= 1u << b;
            dense[(size_t)v * NJ + b] += w;
        }
    auto pop = [](unsigned x) { return __builtin_popcount(x); };
    int nblk = 0, ne = 0;
    for (int seg = 0; seg < SK_NSEG; ++seg) {
        const int v0 = seg * SK_SEG;
        const int n = (NV - v0) < SK_SEG ? (NV - v0) : SK_SEG;
        unsigned best_sel = 0;
        if (n > SK_BV) {
            // choose which n - 8 (or 8) vertices go to the second block; vertex 0 stays in the first
            int best_cost = 1 << 30;
            const int n1 = SK_BV;                              // first block is full, the second takes the rest
            for (unsigned sel = 1; sel < (1u << n); sel += 2) {   // bit i = vertex i in the first block; bit 0 fixed
                if (pop(sel) != n1) continue;
                unsigned ua = 0, ub = 0;
                for (int i = 0; i < n; ++i) ((sel >> i) & 1 ? ua : ub) |= mask[v0 + i];
                const int cost = pop(ua) + pop(ub);
                if (cost < best_cost) { best_cost = cost; best_sel = sel; }
            }
        } else {
            best_sel = (1u << n) - 1;
        }
        for (int half = 0; half < 2; ++half) {
            int grp[SK_BV];
            int m = 0;
            unsigned u = 0;
            for (int i = 0; i < n; ++i)
                if ((int)((best_sel >> i) & 1) == (half == 0 ? 1 : 0)) { grp[m++] = v0 + i; u |= mask[v0 + i]; }
            if (m > SK_BV || nblk >= SK_NBLK) return MB_E_MODEL;
            blk_ptr[nblk] = ne;
            for (int j = 0; j < SK_BV; ++j) {
                perm[nblk * SK_BV + j] = j < m ? grp[j] : -1;
                vloc[nblk * SK_BV + j] = j < m ? (uint8_t)(grp[j] - v0) : (uint8_t)255;
            }
            for (int k = 0; k < NJ; ++k) {
                if (!(u & (1u << k))) continue;
                if (ne >= SK_MAX_ENT) return MB_E_MODEL;
                ent_bone[ne] = k;
                for (int j = 0; j < SK_BV; ++j) ent_w[ne * SK_BV + j] = j < m ? dense[(size_t)grp[j] * NJ + k] : 0.f;
                ++ne;
            }
            if (ne == blk_ptr[nblk]) return MB_E_MODEL;       // a block without any bone (vertex without weights)
            ++nblk;
        }
    }
    if (nblk != SK_NBLK) return MB_E_MODEL;
    blk_ptr[SK_NBLK] = ne;

    // ---- slot schedule (Belady) ----
    {
        std::vector<int> next_use(ne), last_seen(NJ, 1 << 30);
        for (int e = ne - 1; e >= 0; --e) { next_use[e] = last_seen[ent_bone[e]]; last_seen[ent_bone[e]] = e; }
        int slot_bone[SK_SLOTS], slot_next[SK_SLOTS], slot_last[SK_SLOTS];   // occupant, its next use, its last use so far
        for (int s = 0; s < SK_SLOTS; ++s) { slot_bone[s] = -1; slot_next[s] = 1 << 30; slot_last[s] = -1; }
        struct Cmd { int after, slot, bone, next_group; };
        std::vector<Cmd> cmds;
        std::vector<int> first_load_of_slot(SK_SLOTS, -1);
        for (int e = 0; e < ne; ++e) {
            const int k = ent_bone[e];
            int s = -1;
            for (int t = 0; t < SK_SLOTS; ++t) if (slot_bone[t] == k) s = t;
            int wait = 0;
            if (s < 0) {
                for (int t = 0; t < SK_SLOTS && s < 0; ++t) if (slot_bone[t] < 0) s = t;      // an empty slot first
                if (s < 0) { s = 0; for (int t = 1; t < SK_SLOTS; ++t) if (slot_next[t] > slot_next[s]) s = t; }
                Cmd c = {slot_last[s], s, k, 0};
                if (slot_bone[s] < 0) { c.after = -2; c.next_group = 1; first_load_of_slot[s] = (int)cmds.size(); }
                cmds.push_back(c);
                slot_bone[s] = k;
                wait = 1;
            }
            slot_next[s] = next_use[e];
            slot_last[s] = e;
            ent_bone[e] = k | (s << 4) | (wait << 7);
        }
        // the first occupant of a slot is loaded for the NEXT group once this group is done with the slot
        for (int s = 0; s < SK_SLOTS; ++s)
            if (first_load_of_slot[s] >= 0) cmds[first_load_of_slot[s]].after = slot_last[s];
        std::stable_sort(cmds.begin(), cmds.end(), [](const Cmd& a, const Cmd& b) { return a.after < b.after; });
        if ((int)cmds.size() > SK_MAX_CMD) return MB_E_MODEL;
        cmd[0] = (int)cmds.size();
        for (size_t i = 0; i < cmds.size(); ++i) {
            if (cmds[i].after < 0 || cmds[i].after >= 1023) return MB_E_MODEL;
            cmd[1 + i] = (cmds[i].after + 1) | (cmds[i].slot << 10) | (cmds[i].bone << 13) | (cmds[i].next_group << 17);
        }
        // slot contents and command cursor at the start of every segment: a split sweep (small batches) can
        // start there; [7] = the bones the segment's entries read
        int* split = reinterpret_cast<int*>(out + L.sk_split);
        static_assert(SK_SLOTS <= 6, "6 slots + cursor + mask per row");
        for (int p = 0; p < SK_NSEG; ++p) {
            const int e0 = blk_ptr[p * SK_SEG_BLKS];
            int state[SK_SLOTS];
            for (int sl = 0; sl < SK_SLOTS; ++sl) state[sl] = -1;
            for (const Cmd& c : cmds) if (c.next_group) state[c.slot] = c.bone;
            int ci0 = 0;
            for (const Cmd& c : cmds) {
                if (c.after >= e0) break;
                if (!c.next_group) state[c.slot] = c.bone;
                ++ci0;
            }
            for (int sl = 0; sl < 6; ++sl) split[p * 8 + sl] = sl < SK_SLOTS ? state[sl] : -1;
            split[p * 8 + 6] = ci0;
            int touched = 0;
            for (int e = e0; e < blk_ptr[(p + 1) * SK_SEG_BLKS]; ++e) touched |= 1 << (ent_bone[e] & 15);
            split[p * 8 + 7] = touched;
        }
    }
    for (int c = 0; c < SK_TMPL_PAD; ++c) {
        const int p = c / 3;
        coord_map[c] = (p < SK_NPOS && perm[p] >= 0) ? perm[p] * 3 + c % 3 : -1;
    }
    return 0;
}

// Host-side verification of a packed skin program (used by tests/ and by mb_mano_pack_constants itself):
// replays the slot schedule and checks that every entry reads the bone it expects from a slot whose
// load was issued earlier, that loads only replace occupants that are not used again before, and
// that the blocks partition the vertices.  stats[0..3] = entries, commands (loads per sweep), blocks, max bones per block.
int skin_program_check(const void* host_blob, int32_t* stats) {
    const BlobLayout L = blob_layout();
    const char* in = reinterpret_cast<const char*>(host_blob);
    const int* blk_ptr = reinterpret_cast<const int*>(in + L.sk_blk_ptr);
    const int* ent = reinterpret_cast<const int*>(in + L.sk_ent_bone);
    const float* ent_w = reinterpret_cast<const float*>(in + L.sk_ent_w);
    const uint8_t* vloc = reinterpret_cast<const uint8_t*>(in + L.sk_vloc);
    const int* perm = reinterpret_cast<const int*>(in + L.sk_perm);
    const int* cmd = reinterpret_cast<const int*>(in + L.sk_cmd);
    const int ne = blk_ptr[SK_NBLK], ncmd = cmd[0];
    if (ne <= 0 || ne > SK_MAX_ENT || ncmd <= 0 || ncmd > SK_MAX_CMD) return MB_E_MODEL;
    // blocks: a permutation of the vertices, 8 per block, inside their 16-vertex segment
    std::vector<int> seen(NV, 0);
    int max_bones = 0;
    for (int b = 0; b < SK_NBLK; ++b) {
        if (blk_ptr[b + 1] <= blk_ptr[b]) return MB_E_MODEL;
        max_bones = std::max(max_bones, blk_ptr[b + 1] - blk_ptr[b]);
        for (int j = 0; j < SK_BV; ++j) {
            const int v = perm[b * SK_BV + j];
            if (v < 0) { if (vloc[b * SK_BV + j] != 255) return MB_E_MODEL; continue; }
            if (v >= NV || seen[v]++ || v / SK_SEG != b / SK_SEG_BLKS || vloc[b * SK_BV + j] != v % SK_SEG) return MB_E_MODEL;
        }
        for (int e = blk_ptr[b]; e < blk_ptr[b + 1]; ++e)
            for (int j = 0; j < SK_BV; ++j)
                if (perm[b * SK_BV + j] < 0 && ent_w[e * SK_BV + j] != 0.f) return MB_E_MODEL;
    }
    for (int v = 0; v < NV; ++v) if (!seen[v]) return MB_E_MODEL;
    // slot schedule: replay two consecutive groups (the second one starts from the prefetched state)
    int slot_bone[SK_SLOTS], slot_group[SK_SLOTS];
    bool slot_waited[SK_SLOTS];
    for (int s = 0; s < SK_SLOTS; ++s) { slot_bone[s] = -1; slot_group[s] = -1; slot_waited[s] = true; }
    auto issue = [&](int c, int group) -> bool {
        const int s = (c >> 10) & 7, k = (c >> 13) & 15;
        if (s >= SK_SLOTS || !slot_waited[s]) return false;    // previous load of the slot never consumed
        slot_bone[s] = k; slot_group[s] = group; slot_waited[s] = false;
        return true;
    };
    for (int i = 0; i < ncmd; ++i)
        if ((cmd[1 + i] >> 17) & 1) { if (!issue(cmd[1 + i], 0)) return MB_E_MODEL; }
    for (int group = 0; group < 2; ++group) {
        int ci = 0;
        for (int e = 0; e < ne; ++e) {
            const int k = ent[e] & 15, s = (ent[e] >> 4) & 7, wait = (ent[e] >> 7) & 1;
            if (s >= SK_SLOTS || slot_bone[s] != k || slot_group[s] != group) return MB_E_MODEL;
            if (wait) { if (slot_waited[s]) return MB_E_MODEL; slot_waited[s] = true; }
            else if (!slot_waited[s]) return MB_E_MODEL;       // reading a slot whose load was never waited for
            while (ci < ncmd && (cmd[1 + ci] & 1023) == e + 1) {
                const int c = cmd[1 + ci];
                if (!issue(c, group + ((c >> 17) & 1))) return MB_E_MODEL;
                ++ci;
            }
            if (ci < ncmd && (cmd[1 + ci] & 1023) < e + 1) return MB_E_MODEL;   // commands must be sorted
        }
        if (ci != ncmd) return MB_E_MODEL;
    }
    // split sweeps: a sweep started at any segment from its snapshot, without any next-group load, finds its bones
    const int* split = reinterpret_cast<const int*>(in + L.sk_split);
    for (int p = 0; p < SK_NSEG; ++p) {
        const int e0 = blk_ptr[p * SK_SEG_BLKS], e1 = ne;
        int touched = 0;
        for (int e = e0; e < blk_ptr[(p + 1) * SK_SEG_BLKS]; ++e) touched |= 1 << (ent[e] & 15);
        if (split[p * 8 + 7] != touched) return MB_E_MODEL;
        int state[SK_SLOTS];
        for (int s = 0; s < SK_SLOTS; ++s) state[s] = split[p * 8 + s];
        int ci = split[p * 8 + 6];
        for (int e = e0; e < e1; ++e) {
            const int k = ent[e] & 15, s = (ent[e] >> 4) & 7;
            if (state[s] != k) return MB_E_MODEL;
            while (ci < ncmd && (cmd[1 + ci] & 1023) == e + 1) {
                const int c = cmd[1 + ci];
                if (!((c >> 17) & 1)) state[(c >> 10) & 7] = (c >> 13) & 15;
                ++ci;
            }
        }
    }
    if (stats) { stats[0] = ne; stats[1] = ncmd; stats[2] = SK_NBLK; stats[3] = max_bones; }
    return 0;
}

// ---------------------------------------------------------------- launchers
int launch_skin_forward(const void* blob, const float* v_posed_t, const float* bone_t, int B,
                        float* verts, float* joints, cudaStream_t s) {
    if (B <= 0) return 0;
    static SmemAttrOnce once_a, once_b;
    if (int arc = ensure_dyn_smem(once_a, skin_forward_kernel<false>, SKF_SMEM)) return arc;
    if (int arc = ensure_dyn_smem(once_b, skin_forward_kernel<true>, SKF_SMEM)) return arc;
    const int ngroups = (B + 31) >> 5;
    const int spu = skin_segments_per_unit(ngroups, SKF_SWEEPERS);
    if (spu < SK_NSEG) {
        const int nunits = ngroups * skin_units_per_group(spu);
        skin_forward_kernel<true><<<nunits < NUM_SMS ? nunits : NUM_SMS, SKF_THREADS, SKF_SMEM, s>>>(blob, v_posed_t, bone_t, B, verts, joints, spu);
    } else {
        skin_forward_kernel<false><<<NUM_SMS, SKF_THREADS, SKF_SMEM, s>>>(blob, v_posed_t, bone_t, B, verts, joints, SK_NSEG);
    }
    return cuda_rc();
}

int launch_skin_backward(const void* blob, const float* v_posed_t, const float* bone_t, const float* g_verts,
                         const float* g_joints, int B, float* dv_t, unsigned char* dvp, float* dbone, int dbone_hand_minor,
                         float* dparts, cudaStream_t s) {
    if (B <= 0) return 0;
    static SmemAttrOnce once_a, once_b;
    if (int arc = ensure_dyn_smem(once_a, skin_backward_kernel<false>, SKB_SMEM)) return arc;
    if (int arc = ensure_dyn_smem(once_b, skin_backward_kernel<true>, SKB_SMEM)) return arc;
    const int ngroups = (B + 31) >> 5;
    const int spu = skin_segments_per_unit(ngroups, SKB_SWEEPERS);
    static const int role_flip = (getenv("MANO_B200_SKB_FLIP") ? atoi(getenv("MANO_B200_SKB_FLIP")) : 1) |
                                 ((getenv("MANO_B200_SKB_EXP") ? atoi(getenv("MANO_B200_SKB_EXP")) : 0) << 8);
    if (spu < SK_NSEG) {
        if (dparts == nullptr) return MB_E_NULL;
        const int nunits = ngroups * skin_units_per_group(spu);
        skin_backward_kernel<true><<<nunits < NUM_SMS ? nunits : NUM_SMS, SKB_THREADS, SKB_SMEM, s>>>(
            blob, v_posed_t, bone_t, g_verts, g_joints, B, dv_t, dvp, dbone, dbone_hand_minor, dparts, spu, role_flip);
        int rc = cuda_rc();
        if (rc) return rc;
        dbone_reduce_kernel<<<ngroups * NJ, BONE_F * 32, 0, s>>>(blob, dparts, B, spu, dbone, dbone_hand_minor);
    } else {
        skin_backward_kernel<false><<<NUM_SMS, SKB_THREADS, SKB_SMEM, s>>>(blob, v_posed_t, bone_t, g_verts, g_joints, B, dv_t, dvp,
                                                                          dbone, dbone_hand_minor, nullptr, SK_NSEG, role_flip);
    }
    return cuda_rc();
}

int launch_rows_to_t(const void* blob, const float* rows, int pitch, int B, float* t, cudaStream_t s) {
    if (B <= 0) return 0;
    const long long n = (long long)((B + 31) >> 5) * SK_NCOORD * 32;
    rows_to_t_kernel<<<conv_grid(n), 256, 0, s>>>(blob, rows, pitch, B, t);
    return cuda_rc();
}
int launch_t_to_rows(const void* blob, const float* t, int pitch, int B, float* rows, cudaStream_t s) {
    if (B <= 0) return 0;
    const long long n = (long long)((B + 31) >> 5) * SK_NCOORD * 32;
    t_to_rows_kernel<<<conv_grid(n), 256, 0, s>>>(blob, t, pitch, B, rows);
    return cuda_rc();
}
int launch_bone_rows_to_t(const float* bone, int B, float* bone_t, cudaStream_t s) {
    if (B <= 0) return 0;
    const long long n = (long long)((B + 31) >> 5) * GROUP_BONE_FLOATS;
    bone_rows_to_t_kernel<<<conv_grid(n), 256, 0, s>>>(bone, B, bone_t);
    return cuda_rc();
}

}  // namespace mb
