// mano_lbs.cu — linear blend skinning over 778 vertices x 16 bones, forward and backward.
//
// Reference: MANOLayer.py:177-185 (T = sum_k w_vk A_k ; v' = T [v_posed;1]), :190-202 (the five
// fingertip vertices become joints 4,8,12,16,20) and :188,:204-205 (global rotation — already
// folded into the bone transforms by the pose stage, so it costs nothing here).
//
// Mapping (round-1 ncu: the vertex-per-lane version spent 5.6 k warp-instructions per hand on
// divergent per-vertex bone loops): LANE = HAND.  A CTA owns a group of 32 hands and sweeps the
// 778 vertices in 13 chunks of 64.  A chunk is staged in shared memory TRANSPOSED — tile[f][hand]
// with pitch 33 — so that
//   * global traffic is coalesced: each hand row contributes contiguous 768-byte segments, read as
//     16-byte (v_posed) / 8-byte (g_verts, whose rows are only 8-byte aligned) vectors and written
//     back as 16-/8-byte vectors;
//   * every shared-memory access of the compute phase is conflict-free (lane = hand = bank);
//   * the skinning weights / bone ids of a vertex are WARP-UNIFORM: no divergence, weight loads
//     are broadcasts, and the per-vertex bone loop costs exactly its nnz.
// The 16 bone transforms of the 32 hands sit in shared memory with pitch 196 floats (49 x 16 B,
// odd) so the three float4 rows of a bone are conflict-free LDS.128.
// Loads of chunk i+1 are issued into registers before chunk i is computed (software pipeline).
//
// Backward:  dv_posed_v = sum_k w_vk R'_k^T g_v  and
// dA'_k = sum_v w_vk g_v (x) [v_posed_v ; 1].  The second is a reduction over vertices: each warp
// OWNS up to three bones (or halves of long bones; host-side greedy balance) and keeps their 3x4
// accumulators in registers across the whole vertex sweep — no atomics, no shuffles.
#include <cuda_bf16.h>
#include "common.cuh"
#include "blend_tc.cuh"

namespace mb {
namespace {

constexpr int LBS_THREADS = LBS_WARPS * 32;     // 512
constexpr int TP = 33;                          // transposed tile pitch (floats)
constexpr int BONE_PITCH = 196;                 // floats per hand in the bone tile (49 float4, odd)
constexpr int HG = 32;                          // hands per group
__constant__ int c_tip_vert[5] = {333, 444, 672, 555, 745};
__constant__ int c_tip_slot[5] = {4, 8, 12, 16, 20};

__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float2 ld_stream2(const float* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream4(float* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream2(float* p, const float2& v) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(v.x), "f"(v.y) : "memory");
}

struct SkinCsr {            // per-vertex skinning lists, shared by forward and backward
    int ptr[NV + 1];
    float w[MAX_NNZ];
    uint8_t b[MAX_NNZ];
};

__device__ __forceinline__ void stage_csr(SkinCsr& S, const void* blob) {
    const BlobLayout L = blob_layout();
    const int* p = blob_ptr<int>(blob, L.csr_ptr);
    const float* w = blob_ptr<float>(blob, L.csr_w);
    const uint8_t* b = blob_ptr<uint8_t>(blob, L.csr_b);
    for (int i = threadIdx.x; i <= NV; i += blockDim.x) S.ptr[i] = p[i];
    const int nnz = p[NV];
    for (int i = threadIdx.x; i < nnz; i += blockDim.x) { S.w[i] = w[i]; S.b[i] = b[i]; }
}

// coalesced copy of the 16 bone transforms of `nh` hands into the padded bone tile
__device__ __forceinline__ void stage_bones(float* s_bone, const float* __restrict__ bone, long long h0, int nh) {
    const float4* src = reinterpret_cast<const float4*>(bone + h0 * (NJ * BONE_F));
    for (int i = threadIdx.x; i < nh * (NJ * BONE_F / 4); i += blockDim.x) {
        const int h = i / (NJ * BONE_F / 4), q = i - h * (NJ * BONE_F / 4);
        reinterpret_cast<float4*>(s_bone + h * BONE_PITCH)[q] = src[i];
    }
}

// Register-staged chunk I/O.  A chunk row holds PPR pieces of VEC floats (16- or 8-byte vectors).
// One warp instruction covers LH rows x LQ pieces with LQ * VEC * 4 = 128 contiguous bytes per row:
//   global side : LH fully used 128-byte lines per request;
//   shared side : tile[(q*VEC + e) * 33 + h] -> bank (VEC*q + e + h) mod 32, distinct for the LQ x LH
//                 lanes, so the transposing STS/LDS are conflict-free.
// The (row, piece) of every thread and iteration is a compile-time function of (warp, lane, i).
template <int VEC, int PPR>
struct ChunkMap {
    static constexpr int LQ = 32 / VEC;                 // pieces per row per instruction (8 or 16)
    static constexpr int LH = 32 / LQ;                  // rows per instruction (4 or 2)
    static constexpr int QB = (PPR + LQ - 1) / LQ;      // piece blocks per row
    static constexpr int HB = HG / LH;                  // row blocks
    static constexpr int ITERS = (QB * HB + LBS_WARPS - 1) / LBS_WARPS;
    __device__ __forceinline__ static bool map(int i, int nh, int& h, int& q) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int t = warp + LBS_WARPS * i;
        q = (t % QB) * LQ + (lane & (LQ - 1));
        h = (t / QB) * LH + (lane / LQ);
        return t < QB * HB && q < PPR && h < nh;
    }
};
template <int VEC, int ITERS>
struct ChunkRegs { float v[ITERS][VEC]; };

template <int VEC, int PPR, int MAXI>
__device__ __forceinline__ void chunk_load(ChunkRegs<VEC, MAXI>& R, const float* __restrict__ base, long long pitch,
                                           long long h0, int nh, int f0) {
    using M = ChunkMap<VEC, PPR>;
#pragma unroll
    for (int i = 0; i < M::ITERS; ++i) {
        int h, q;
        if (M::map(i, nh, h, q)) {
            const float* src = base + (h0 + h) * pitch + f0 + q * VEC;
            if (VEC == 4) { float4 t = ld_stream4(src); R.v[i][0] = t.x; R.v[i][1] = t.y; R.v[i][2] = t.z; R.v[i][3] = t.w; }
            else          { float2 t = ld_stream2(src); R.v[i][0] = t.x; R.v[i][1] = t.y; }
        }
    }
}
template <int VEC, int PPR, int MAXI>
__device__ __forceinline__ void chunk_to_tile(const ChunkRegs<VEC, MAXI>& R, float* tile, int nh) {
    using M = ChunkMap<VEC, PPR>;
#pragma unroll
    for (int i = 0; i < M::ITERS; ++i) {
        int h, q;
        if (M::map(i, nh, h, q)) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) tile[(q * VEC + e) * TP + h] = R.v[i][e];
        }
    }
}
// coalesced store of a transposed tile back to rows (PPR pieces of VEC floats per row)
template <int VEC, int PPR>
__device__ __forceinline__ void tile_store(const float* tile, float* __restrict__ base, long long pitch, long long h0,
                                           int nh, int f0) {
    using M = ChunkMap<VEC, PPR>;
#pragma unroll
    for (int i = 0; i < M::ITERS; ++i) {
        int h, q;
        if (M::map(i, nh, h, q)) {
            float* dst = base + (h0 + h) * pitch + f0 + q * VEC;
            if (VEC == 4) st_stream4(dst, make_float4(tile[(q * 4) * TP + h], tile[(q * 4 + 1) * TP + h],
                                                      tile[(q * 4 + 2) * TP + h], tile[(q * 4 + 3) * TP + h]));
            else          st_stream2(dst, make_float2(tile[(q * 2) * TP + h], tile[(q * 2 + 1) * TP + h]));
        }
    }
}

constexpr int TAIL_V = NV - (LBS_CHUNKS - 1) * LBS_CV;       // 10 vertices in the last chunk
constexpr int TAIL_F4 = VP_PITCH - (LBS_CHUNKS - 1) * LBS_CF; // 32 floats of the padded v_posed row
constexpr int TAIL_F2 = TAIL_V * 3;                          // 30 floats of the dense rows
constexpr int I4 = ChunkMap<4, LBS_CF / 4>::ITERS;           // 6
constexpr int I2 = ChunkMap<2, LBS_CF / 2>::ITERS;           // 12

// ------------------------------------------------------------------ forward
struct FwdShared {
    SkinCsr csr;
    alignas(16) float bone[HG * BONE_PITCH];
    alignas(16) float tile[2][LBS_CF * TP];
};

__global__ void __launch_bounds__(LBS_THREADS, 2)
lbs_forward_kernel(const void* __restrict__ blob, const float* __restrict__ v_posed, int pitch,
                   const float* __restrict__ bone, int B, float* __restrict__ verts, float* __restrict__ joints) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FwdShared& S = *reinterpret_cast<FwdShared*>(smem_raw);
    stage_csr(S.csr, blob);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ngroups = (B + HG - 1) / HG;
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const long long h0 = (long long)grp * HG;
        const int nh = (B - h0) < HG ? (int)(B - h0) : HG;
        __syncthreads();                                   // previous group fully stored; csr staged
        stage_bones(S.bone, bone, h0, nh);
        ChunkRegs<4, I4> regs;
        chunk_load<4, LBS_CF / 4, I4>(regs, v_posed, pitch, h0, nh, 0);
        for (int c = 0; c < LBS_CHUNKS; ++c) {
            float* tile = S.tile[c & 1];
            const int f0 = c * LBS_CF;
            const bool last = (c == LBS_CHUNKS - 1);
            const int nv = last ? TAIL_V : LBS_CV;
            if (!last) chunk_to_tile<4, LBS_CF / 4, I4>(regs, tile, nh);
            else       chunk_to_tile<4, TAIL_F4 / 4, I4>(regs, tile, nh);
            if (c + 2 < LBS_CHUNKS)       chunk_load<4, LBS_CF / 4, I4>(regs, v_posed, pitch, h0, nh, f0 + LBS_CF);
            else if (c + 2 == LBS_CHUNKS) chunk_load<4, TAIL_F4 / 4, I4>(regs, v_posed, pitch, h0, nh, f0 + LBS_CF);
            __syncthreads();                               // tile (and, for c == 0, bones) visible
            const float* A0 = S.bone + lane * BONE_PITCH;
            for (int vl = warp; vl < nv; vl += LBS_WARPS) {
                const int v = c * LBS_CV + vl;
                const float x = tile[(vl * 3) * TP + lane], y = tile[(vl * 3 + 1) * TP + lane], z = tile[(vl * 3 + 2) * TP + lane];
                float ox = 0.f, oy = 0.f, oz = 0.f;
                const int e1 = S.csr.ptr[v + 1];
                for (int e = S.csr.ptr[v]; e < e1; ++e) {
                    const float w = S.csr.w[e];
                    const float4* A = reinterpret_cast<const float4*>(A0 + S.csr.b[e] * BONE_F);
                    const float4 r0 = A[0], r1 = A[1], r2 = A[2];
                    ox = fmaf(w, fmaf(r0.x, x, fmaf(r0.y, y, fmaf(r0.z, z, r0.w))), ox);
                    oy = fmaf(w, fmaf(r1.x, x, fmaf(r1.y, y, fmaf(r1.z, z, r1.w))), oy);
                    oz = fmaf(w, fmaf(r2.x, x, fmaf(r2.y, y, fmaf(r2.z, z, r2.w))), oz);
                }
                tile[(vl * 3) * TP + lane] = ox; tile[(vl * 3 + 1) * TP + lane] = oy; tile[(vl * 3 + 2) * TP + lane] = oz;
            }
            __syncthreads();                               // results complete
            if (!last) tile_store<2, LBS_CF / 2>(tile, verts, NVC, h0, nh, f0);
            else       tile_store<2, TAIL_F2 / 2>(tile, verts, NVC, h0, nh, f0);
            if (joints != nullptr && warp < 5 && c_tip_vert[warp] / LBS_CV == c && lane < nh) {
                const int vl = c_tip_vert[warp] - c * LBS_CV;
                float* o = joints + (h0 + lane) * (NOUTJ * 3) + c_tip_slot[warp] * 3;
                o[0] = tile[(vl * 3) * TP + lane]; o[1] = tile[(vl * 3 + 1) * TP + lane]; o[2] = tile[(vl * 3 + 2) * TP + lane];
            }
            // tile (c & 1) is rewritten at chunk c + 2, after the two barriers of chunk c + 1
        }
    }
}

// ----------------------------------------------------------------- backward
struct BwdShared {
    SkinCsr csr;
    int seg[LBS_WARPS][LBS_CHUNKS][LBS_SLOTS][2];
    int slot[LBS_WARPS][LBS_SLOTS];
    unsigned short ent_idx[MAX_NNZ];
    float ent_w[MAX_NNZ];
    alignas(16) float bone[HG * BONE_PITCH];
    alignas(16) float tile_g[LBS_CF * TP];          // single-buffered: chunk c+1 waits in registers
    alignas(16) float tile_v[LBS_CF * TP];
    alignas(16) float tile_o[LBS_CF * TP];          // dv chunk
};
static_assert(sizeof(float) * LBS_WARPS * LBS_SLOTS * BONE_F * (HG + 1) <= sizeof(float) * 2 * LBS_CF * TP,
              "per-slot partial sums are staged in the (idle) g tiles at group end");

__global__ void __launch_bounds__(LBS_THREADS)
lbs_backward_kernel(const void* __restrict__ blob, const float* v_posed, int pitch,
                    const float* __restrict__ bone, const float* __restrict__ g_verts,
                    const float* __restrict__ g_joints, int B,
                    float* dv_posed, unsigned char* __restrict__ dvp,
                    float* __restrict__ dbone) {   // dv_posed may alias v_posed (chunk read before write)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdShared& S = *reinterpret_cast<BwdShared*>(smem_raw);
    stage_csr(S.csr, blob);
    {
        const BlobLayout L = blob_layout();
        const int* seg = blob_ptr<int>(blob, L.bseg);
        const int* slot = blob_ptr<int>(blob, L.bslot);
        const unsigned short* ei = blob_ptr<unsigned short>(blob, L.bent_idx);
        const float* ew = blob_ptr<float>(blob, L.bent_w);
        for (int i = threadIdx.x; i < LBS_WARPS * LBS_CHUNKS * LBS_SLOTS * 2; i += blockDim.x) (&S.seg[0][0][0][0])[i] = seg[i];
        for (int i = threadIdx.x; i < LBS_WARPS * LBS_SLOTS; i += blockDim.x) (&S.slot[0][0])[i] = slot[i];
        const int nnz = blob_ptr<int>(blob, L.csr_ptr)[NV];
        for (int i = threadIdx.x; i < nnz; i += blockDim.x) { S.ent_idx[i] = ei[i]; S.ent_w[i] = ew[i]; }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ngroups = (B + HG - 1) / HG;
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const long long h0 = (long long)grp * HG;
        const int nh = (B - h0) < HG ? (int)(B - h0) : HG;
        __syncthreads();
        stage_bones(S.bone, bone, h0, nh);
        float acc[LBS_SLOTS][BONE_F];
#pragma unroll
        for (int s = 0; s < LBS_SLOTS; ++s)
#pragma unroll
            for (int e = 0; e < BONE_F; ++e) acc[s][e] = 0.f;
        ChunkRegs<4, I4> rv;
        ChunkRegs<2, I2> rg;
        chunk_load<4, LBS_CF / 4, I4>(rv, v_posed, pitch, h0, nh, 0);
        chunk_load<2, LBS_CF / 2, I2>(rg, g_verts, NVC, h0, nh, 0);
        for (int c = 0; c < LBS_CHUNKS; ++c) {
            float* tg = S.tile_g;
            float* tv = S.tile_v;
            const int f0 = c * LBS_CF;
            const bool last = (c == LBS_CHUNKS - 1);
            const int nv = last ? TAIL_V : LBS_CV;
            if (!last) { chunk_to_tile<4, LBS_CF / 4, I4>(rv, tv, nh); chunk_to_tile<2, LBS_CF / 2, I2>(rg, tg, nh); }
            else       { chunk_to_tile<4, TAIL_F4 / 4, I4>(rv, tv, nh); chunk_to_tile<2, TAIL_F2 / 2, I2>(rg, tg, nh); }
            __syncthreads();                               // tiles of chunk c complete (and dv of chunk c-1 stored)
            // fingertip joints are vertices: their upstream gradient joins g_verts (A.2 step 1)
            if (warp < 5 && c_tip_vert[warp] / LBS_CV == c && lane < nh) {
                const int vl = c_tip_vert[warp] - c * LBS_CV;
                const float* gj = g_joints + (h0 + lane) * (NOUTJ * 3) + c_tip_slot[warp] * 3;
                tg[(vl * 3) * TP + lane] += gj[0]; tg[(vl * 3 + 1) * TP + lane] += gj[1]; tg[(vl * 3 + 2) * TP + lane] += gj[2];
            }
            // prefetch chunk c+1 into registers (its v_posed floats are read before dv chunk c+1 overwrites them)
            if (c + 2 < LBS_CHUNKS) {
                chunk_load<4, LBS_CF / 4, I4>(rv, v_posed, pitch, h0, nh, f0 + LBS_CF);
                chunk_load<2, LBS_CF / 2, I2>(rg, g_verts, NVC, h0, nh, f0 + LBS_CF);
            } else if (c + 2 == LBS_CHUNKS) {
                chunk_load<4, TAIL_F4 / 4, I4>(rv, v_posed, pitch, h0, nh, f0 + LBS_CF);
                chunk_load<2, TAIL_F2 / 2, I2>(rg, g_verts, NVC, h0, nh, f0 + LBS_CF);
            }
            __syncthreads();                               // tip fix-up visible
            // (a) dv = sum_s w R'^T g  for this warp's vertices of the chunk
            const float* A0 = S.bone + lane * BONE_PITCH;
            for (int vl = warp; vl < nv; vl += LBS_WARPS) {
                const int v = c * LBS_CV + vl;
                const float gx = tg[(vl * 3) * TP + lane], gy = tg[(vl * 3 + 1) * TP + lane], gz = tg[(vl * 3 + 2) * TP + lane];
                float dx = 0.f, dy = 0.f, dz = 0.f;
                const int e1 = S.csr.ptr[v + 1];
                for (int e = S.csr.ptr[v]; e < e1; ++e) {
                    const float w = S.csr.w[e];
                    const float4* A = reinterpret_cast<const float4*>(A0 + S.csr.b[e] * BONE_F);
                    const float4 r0 = A[0], r1 = A[1], r2 = A[2];
                    const float wx = w * gx, wy = w * gy, wz = w * gz;
                    dx = fmaf(r0.x, wx, fmaf(r1.x, wy, fmaf(r2.x, wz, dx)));
                    dy = fmaf(r0.y, wx, fmaf(r1.y, wy, fmaf(r2.y, wz, dy)));
                    dz = fmaf(r0.z, wx, fmaf(r1.z, wy, fmaf(r2.z, wz, dz)));
                }
                S.tile_o[(vl * 3) * TP + lane] = dx; S.tile_o[(vl * 3 + 1) * TP + lane] = dy; S.tile_o[(vl * 3 + 2) * TP + lane] = dz;
            }
            if (last && threadIdx.x < (VP_PITCH - NVC) * HG)    // zero the two pad floats of the row
                S.tile_o[(TAIL_F2 + threadIdx.x / HG) * TP + (threadIdx.x % HG)] = 0.f;
            // (b) per-bone sums for the slots this warp owns: registers across the whole sweep
#pragma unroll
            for (int s = 0; s < LBS_SLOTS; ++s) {
                const int e1 = S.seg[warp][c][s][1];
                for (int e = S.seg[warp][c][s][0]; e < e1; ++e) {
                    const int fi = S.ent_idx[e];
                    const float w = S.ent_w[e];
                    const float wx = w * tg[fi * TP + lane], wy = w * tg[(fi + 1) * TP + lane], wz = w * tg[(fi + 2) * TP + lane];
                    const float x = tv[fi * TP + lane], y = tv[(fi + 1) * TP + lane], z = tv[(fi + 2) * TP + lane];
                    acc[s][0] = fmaf(wx, x, acc[s][0]); acc[s][1] = fmaf(wx, y, acc[s][1]); acc[s][2] = fmaf(wx, z, acc[s][2]);   acc[s][3] += wx;
                    acc[s][4] = fmaf(wy, x, acc[s][4]); acc[s][5] = fmaf(wy, y, acc[s][5]); acc[s][6] = fmaf(wy, z, acc[s][6]);   acc[s][7] += wy;
                    acc[s][8] = fmaf(wz, x, acc[s][8]); acc[s][9] = fmaf(wz, y, acc[s][9]); acc[s][10] = fmaf(wz, z, acc[s][10]); acc[s][11] += wz;
                }
            }
            __syncthreads();                               // dv chunk complete; g / v tiles free for chunk c+1
            if (dvp == nullptr) {
                if (!last) tile_store<4, LBS_CF / 4>(S.tile_o, dv_posed, pitch, h0, nh, f0);
                else       tile_store<4, TAIL_F4 / 4>(S.tile_o, dv_posed, pitch, h0, nh, f0);
            } else {
                // A operand of the tcgen05 gradient contraction: bf16 hi + mid, UMMA canonical K-major
                // blocks.  One warp store = 32 lanes x 16 B = one contiguous 512-byte piece
                // (4 k-groups x 8 hands of one 8-row group); lane = (k-group, hand) is conflict-free
                // on the transposed tile.
                const int nblk = last ? 1 : LBS_CF / TC_K_CHUNK;           // K chunks of 32 in this vertex chunk
                const int kg = lane >> 3, r = lane & 7;
                unsigned char* tile_base = dvp + (size_t)(grp >> 2) * TCB_A_TILE_BYTES;
                const int rg0 = (grp & 3) * 4;
                for (int piece = warp; piece < nblk * 4; piece += LBS_WARPS) {
                    const int blk = piece >> 2, rgl = piece & 3;
                    const float* src = S.tile_o + ((blk * 4 + kg) * 8) * TP + rgl * 8 + r;
                    __align__(16) __nv_bfloat16 hi[8], mid[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float x = src[e * TP];
                        hi[e] = __float2bfloat16_rn(x);
                        mid[e] = __float2bfloat16_rn(x - __bfloat162float(hi[e]));
                    }
                    unsigned char* dst = tile_base + (size_t)(c * (LBS_CF / TC_K_CHUNK) + blk) * TCB_A_CHUNK_BYTES +
                                         (((rg0 + rgl) * 4 + kg) * 8 + r) * 16;
                    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
                    *reinterpret_cast<uint4*>(dst + TC_A_BLOCK_BYTES) = *reinterpret_cast<const uint4*>(mid);
                }
            }
        }
        // ---- group end: combine the per-slot sums into dbone[h][k][12] (staged in the now idle g tiles)
        float (*part)[BONE_F][HG + 1] = reinterpret_cast<float (*)[BONE_F][HG + 1]>(&S.tile_g[0]);
#pragma unroll
        for (int s = 0; s < LBS_SLOTS; ++s)
#pragma unroll
            for (int e = 0; e < BONE_F; ++e) part[warp * LBS_SLOTS + s][e][lane] = acc[s][e];
        __syncthreads();
        for (int i = threadIdx.x; i < nh * NJ * BONE_F; i += LBS_THREADS) {
            const int h = i / (NJ * BONE_F), r = i - h * (NJ * BONE_F), k = r / BONE_F, e = r - k * BONE_F;
            float sum = 0.f;
#pragma unroll
            for (int sl = 0; sl < LBS_WARPS * LBS_SLOTS; ++sl) {
                const int code = (&S.slot[0][0])[sl];
                if (code >= 0 && (code & 255) == k) sum += part[sl][e][h];
            }
            dbone[(h0 + h) * (NJ * BONE_F) + r] = sum;
        }
    }
}

}  // namespace

int launch_lbs_forward(const void* blob, const float* v_posed, int pitch, const float* bone, int B,
                       float* verts, float* joints, cudaStream_t s) {
    if (B <= 0) return 0;
    static bool attr_done = false;
    const size_t smem = sizeof(FwdShared);
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(lbs_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    const int ngroups = (B + HG - 1) / HG;
    const int cap = NUM_SMS * 2;
    lbs_forward_kernel<<<ngroups < cap ? ngroups : cap, LBS_THREADS, smem, s>>>(blob, v_posed, pitch, bone, B, verts, joints);
    return cuda_rc();
}

int launch_lbs_backward(const void* blob, const float* v_posed, int pitch, const float* bone,
                        const float* g_verts, const float* g_joints, int B,
                        float* dv_posed, unsigned char* dvp, float* dbone, cudaStream_t s) {
    if (B <= 0) return 0;
    static bool attr_done = false;
    const size_t smem = sizeof(BwdShared);
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(lbs_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    const int ngroups = (B + HG - 1) / HG;
    const int cap = NUM_SMS;
    lbs_backward_kernel<<<ngroups < cap ? ngroups : cap, LBS_THREADS, smem, s>>>(blob, v_posed, pitch, bone, g_verts, g_joints,
                                                                                 B, dv_posed, dvp, dbone);
    return cuda_rc();
}

}  // namespace mb
