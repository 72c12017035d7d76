// blend_tc.cu — the blend-shape contraction on the 5th-generation tensor cores (tcgen05),
// MB_MODE_F16X3 (fp32-accurate) and MB_MODE_F16 (fast).
//
//   v_posed[h][n] = v_template[n] + sum_k f[h][k] * basis[k][n]          (MANOLayer.py:130-137)
//   f = [beta(10) | vec(R_j - I)(135)], k padded to 160;  n = block-order coordinate, 2352 -> 15 tiles of 160
//
// GEMM mapping: M = 128 hands (TMEM lanes), N = 160 vertex coordinates, K = 160.
//   * B operand (the basis slice of the CTA's N-tile, 100 KB as fp16 hi+lo) is RESIDENT in shared
//     memory: it is fetched once per N-tile by TMA bulk copies from a pre-tiled image in the
//     constant blob and reused for every hand tile the CTA processes.
//   * A operand (feature rows, written by the pose stage as fp16 hi/lo in the UMMA canonical
//     K-major layout) streams through a 4-stage TMA/mbarrier ring, 16 KB (hi+lo of a K=32 chunk)
//     per stage.
//   * one elected thread issues tcgen05.mma.kind::f16 (fp32 accumulate in TMEM); two accumulator
//     stages (2 x 160 TMEM columns) let the epilogue of tile i overlap the MMAs of tile i+1.
//   * epilogue warps: tcgen05.ld (lane = hand) -> add v_template in fp32 -> straight from registers
//     into the HAND-MINOR scratch v_posed_t[group][column][32]: a TMEM lane quarter is a hand group,
//     so every column of a warp is one coalesced 128-byte store and no transposition is needed.
//     Columns are in the skinning kernels' block order (the basis image is permuted at pack time).
// fp32 accuracy from fp16 tensor cores: both operands are split x = hi + lo (two fp16, 22
// significand bits) after a power-of-two pre-scale that keeps lo out of the fp16 subnormals, and
// three products hi*hi + lo*hi + hi*lo are accumulated in fp32; v_template (the one large term)
// never enters the tensor core — it is added in the epilogue.  MB_MODE_F16 issues hi*hi only.
// Tiles are distributed as contiguous ranges of the flattened (n_tile, m_tile) space, so every
// CTA gets the same number of tiles (+-1) and reloads its resident B at most once.
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <string.h>
#include <math.h>
#include "common.cuh"
#include "blend_tc.cuh"
#include "ptx.cuh"
#include "tc_ptx.cuh"

namespace mb {
namespace {

// ---------------------------------------------------------------- kernel
constexpr int TC_THREADS = 384;                   // w0 TMA, w1 MMA, w2 TMEM alloc, w3 idle, w4-11 epilogue
constexpr int EPI_WARPS = 8;
constexpr int A_STAGES = TC_K_CHUNKS;             // one ring slot per K chunk: slot index == chunk index (static descriptors)
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = 512;
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);   // f16 x f16 -> f32, K-major A and B

struct TcShared {
    alignas(128) unsigned char b[TC_B_TILE_BYTES];                  // resident basis tile (hi+lo, 5 chunks)
    alignas(128) unsigned char a[A_STAGES][TC_A_STAGE_BYTES];       // feature ring
    alignas(8) unsigned long long full[A_STAGES], empty[A_STAGES];
    unsigned long long acc_full[ACC_STAGES], acc_empty[ACC_STAGES];
    unsigned long long b_full;
    uint32_t tmem_base;
    int abort_flag;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
blend_tc_forward_kernel(const TcBlobHeader* __restrict__ hdr, const unsigned char* __restrict__ basis_tc,
                        const float* __restrict__ tmpl, const unsigned char* __restrict__ featp,
                        float* __restrict__ v_posed_t, int B, int m_tiles, int products) {
    // declared with its alignment and used WITHOUT integer arithmetic on the address: rounding the pointer up through uintptr_t
    // made the compiler lose the shared address space — every access became a generic LD / ST [profiles/r2]
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcShared& S = *reinterpret_cast<TcShared*>(smem_raw);
    const float out_scale = exp2f(-(float)(hdr->basis_scale_log2 + hdr->feat_scale_log2));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // Tile schedule: CTA c owns n-tile (c % 15) for its whole life (its basis slice stays resident in
    // shared memory) and walks the hand tiles m = c/15, c/15 + G, ... where G is the number of CTAs
    // on that n-tile.  All 15 groups sweep m upwards at the same pace, so a hand tile's feature rows
    // are fetched from HBM once and served from L2 to the other 14 groups (round-1 ncu: the n-major
    // order re-read them from DRAM 15 times).
    const int n_tile = blockIdx.x % TC_N_TILES;
    const int m_first = blockIdx.x / TC_N_TILES;
    const int m_step = ((int)gridDim.x - n_tile + TC_N_TILES - 1) / TC_N_TILES;

    if (threadIdx.x == 0) {
        for (int s = 0; s < A_STAGES; ++s) { mbar_init(smem_u32(&S.full[s]), 1); mbar_init(smem_u32(&S.empty[s]), 1); }
        for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(smem_u32(&S.acc_full[s]), 1); mbar_init(smem_u32(&S.acc_empty[s]), EPI_WARPS); }
        mbar_init(smem_u32(&S.b_full), 1);
        S.abort_flag = 0;
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(&S.tmem_base), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    volatile int* abort_flag = &S.abort_flag;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer: resident B once, then feature chunks through the ring =====
        mbar_expect_tx(smem_u32(&S.b_full), TC_B_TILE_BYTES);
        const unsigned char* src = basis_tc + (size_t)n_tile * TC_B_TILE_BYTES;
        for (int i = 0; i < TC_B_TILE_BYTES / TC_B_BLOCK_BYTES; ++i)
            bulk_g2s(smem_u32(S.b) + i * TC_B_BLOCK_BYTES, src + (size_t)i * TC_B_BLOCK_BYTES, TC_B_BLOCK_BYTES, smem_u32(&S.b_full));
        uint32_t phase = 0;
        bool ok = true;
        for (int m_tile = m_first; m_tile < m_tiles && ok; m_tile += m_step) {
            const unsigned char* fsrc = featp + (size_t)m_tile * TC_A_TILE_BYTES;
#pragma unroll
            for (int c = 0; c < TC_K_CHUNKS; ++c) {
                if (!(ok = mbar_wait(smem_u32(&S.empty[c]), phase ^ 1, abort_flag))) break;
                mbar_expect_tx(smem_u32(&S.full[c]), TC_A_STAGE_BYTES);
                bulk_g2s(smem_u32(S.a[c]), fsrc + (size_t)c * TC_A_STAGE_BYTES, TC_A_STAGE_BYTES, smem_u32(&S.full[c]));
            }
            phase ^= 1;
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer: 30 (or 10) fully unrolled tcgen05.mma per tile, descriptors are base + constant =====
        uint32_t phase = 0, acc = 0, acc_phase = 0;
        bool ok = mbar_wait(smem_u32(&S.b_full), 0, abort_flag);
        const uint64_t a_base = umma_desc(smem_u32(S.a[0]), TC_LBO, TC_SBO);
        const uint64_t b_base = umma_desc(smem_u32(S.b), TC_LBO, TC_SBO);
        for (int m_tile = m_first; m_tile < m_tiles && ok; m_tile += m_step) {
            if (!(ok = mbar_wait(smem_u32(&S.acc_empty[acc]), acc_phase ^ 1, abort_flag))) break;
            tc_fence_after();
            const uint32_t d_tmem = tmem + acc * TC_N;
#pragma unroll
            for (int c = 0; c < TC_K_CHUNKS; ++c) {
                if (!(ok = mbar_wait(smem_u32(&S.full[c]), phase, abort_flag))) break;
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < TC_K_CHUNK / 16; ++j) {
                    const uint64_t a_hi = a_base + (uint64_t)((c * TC_A_STAGE_BYTES + j * 2 * (int)TC_LBO) >> 4);
                    const uint64_t a_lo = a_hi + (uint64_t)(TC_A_BLOCK_BYTES >> 4);
                    const uint64_t b_hi = b_base + (uint64_t)((c * 2 * TC_B_BLOCK_BYTES + j * 2 * (int)TC_LBO) >> 4);
                    const uint64_t b_lo = b_hi + (uint64_t)(TC_B_BLOCK_BYTES >> 4);
                    umma_f16(d_tmem, a_hi, b_hi, IDESC, (c | j) ? 1u : 0u);
                    if (products == 3) {
                        umma_f16(d_tmem, a_lo, b_hi, IDESC, 1);
                        umma_f16(d_tmem, a_hi, b_lo, IDESC, 1);
                    }
                }
                tc_commit(smem_u32(&S.empty[c]));              // frees the feature slot when these MMAs retire
            }
            if (!ok) break;
            tc_commit(smem_u32(&S.acc_full[acc]));             // accumulator ready for the epilogue
            phase ^= 1;
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> hand-minor v_posed_t =====
        // warp % 4 selects the TMEM lane quarter (= one group of 32 hands); warps 4-7 take column
        // chunks 0-2, warps 8-11 chunks 3-4
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int j_begin = half == 0 ? 0 : 3, j_end = half == 0 ? 3 : TC_N / 32;
        uint32_t acc = 0, acc_phase = 0;
        bool ok = true;
        const int n0 = n_tile * TC_N;
        for (int m_tile = m_first; m_tile < m_tiles && ok; m_tile += m_step) {
            ok = __all_sync(0xffffffffu, mbar_wait(smem_u32(&S.acc_full[acc]), acc_phase, abort_flag));
            if (!ok) break;
            tc_fence_after();
            const long long group = (long long)m_tile * (TC_M / 32) + q;
            const bool live = group * 32 < B;                   // whole groups beyond the batch are skipped
#pragma unroll 1
            for (int j = j_begin; j < j_end; ++j) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + acc * TC_N + j * 32, v);
                if (j == j_end - 1) {                           // this warp's share is drained: release the accumulator
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&S.acc_empty[acc]));
                }
                const int col0 = n0 + j * 32;
                const float tv = tmpl[col0 + lane];             // v_template of this chunk's columns (block order, zero padded)
                if (live) {
                    float* dst = v_posed_t + ((size_t)group * SK_NCOORD + col0) * 32 + lane;
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (col0 + c < SK_NCOORD)
                            __stcs(dst + c * 32, fmaf(v[c], out_scale, __shfl_sync(0xffffffffu, tv, c)));
                }
            }
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
    }
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, TMEM_COLS);
    if (threadIdx.x == 0 && S.abort_flag) __trap();           // a stalled pipeline is an error, never a silent result
}


// ----------------------------------------------------------------- large batches: hand tile resident, basis streamed
// ncu on the kernel above at 2^20 hands (profiles/r1): tensor pipe 27 %, DRAM 33 %, and the time does not change
// when two of the three products are dropped — its feature ring holds exactly ONE hand tile, so every tile waits
// a full bulk-copy latency (the next tile's chunk can only be requested when this tile's MMAs on that slot retire).
// Here the roles are swapped: the CTA keeps the 80 KB feature rows of its hand tile resident for all 15 n-tiles,
// and the basis (1.5 MB, always an L2 hit, the same stream for every CTA and every hand tile) flows through a
// BS-stage ring of 20 KB K-chunks that never drains at a tile boundary: the producer runs up to BS chunks
// (> one n-tile) ahead.  Feature rows are read once from HBM instead of 15 times from L2.  A second producer
// thread requests the next hand tile's chunk c as soon as the last n-tile's MMAs on it retire.
constexpr int MRES_ACC = 3;                                        // accumulator stages (3 x 160 TMEM columns)
constexpr int BS = 6;                                              // basis ring stages (20 KB each)
struct TcSharedM {
    alignas(128) unsigned char a[TC_K_CHUNKS][TC_A_STAGE_BYTES];    // resident feature rows of the hand tile (hi+lo)
    alignas(128) unsigned char b[BS][2 * TC_B_BLOCK_BYTES];         // basis ring: (n-tile, K-chunk) hi+lo
    alignas(16) float tmpl[SK_TMPL_PAD];                            // v_template, block order
    alignas(8) unsigned long long a_full[TC_K_CHUNKS], a_empty[TC_K_CHUNKS];
    unsigned long long b_full[BS], b_empty[BS];
    unsigned long long acc_full[MRES_ACC], acc_empty[MRES_ACC];
    uint32_t tmem_base;
    int abort_flag;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
blend_tc_forward_mres_kernel(const TcBlobHeader* __restrict__ hdr, const unsigned char* __restrict__ basis_tc,
                             const float* __restrict__ tmpl, const unsigned char* __restrict__ featp,
                             float* __restrict__ v_posed_t, int B, int m_tiles, int products) {
    // declared with its alignment and used WITHOUT integer arithmetic on the address: rounding the pointer up through uintptr_t
    // made the compiler lose the shared address space — every access became a generic LD / ST [profiles/r2]
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcSharedM& S = *reinterpret_cast<TcSharedM*>(smem_raw);
    const float out_scale = exp2f(-(float)(hdr->basis_scale_log2 + hdr->feat_scale_log2));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int c = 0; c < TC_K_CHUNKS; ++c) { mbar_init(smem_u32(&S.a_full[c]), 1); mbar_init(smem_u32(&S.a_empty[c]), 1); }
        for (int st = 0; st < BS; ++st) { mbar_init(smem_u32(&S.b_full[st]), 1); mbar_init(smem_u32(&S.b_empty[st]), 1); }
        for (int st = 0; st < MRES_ACC; ++st) { mbar_init(smem_u32(&S.acc_full[st]), 1); mbar_init(smem_u32(&S.acc_empty[st]), EPI_WARPS); }
        S.abort_flag = 0;
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < SK_TMPL_PAD; i += TC_THREADS) S.tmpl[i] = tmpl[i];
    if (warp == 2) tmem_alloc(smem_u32(&S.tmem_base), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    volatile int* abort_flag = &S.abort_flag;

    if (warp == 0 && lane == 0) {
        // ===== basis producer: the same 75-chunk stream for every hand tile, kept in L2 =====
        const uint64_t keep = l2_policy_evict_last();
        uint32_t stage = 0, phase = 0;
        bool ok = true;
        for (int m_tile = blockIdx.x; m_tile < m_tiles && ok; m_tile += gridDim.x) {
            for (int i = 0; i < TC_N_TILES * TC_K_CHUNKS; ++i) {
                if (!(ok = mbar_wait(smem_u32(&S.b_empty[stage]), phase ^ 1, abort_flag))) break;
                mbar_expect_tx(smem_u32(&S.b_full[stage]), 2 * TC_B_BLOCK_BYTES);
                bulk_g2s_hint(smem_u32(S.b[stage]), basis_tc + (size_t)i * 2 * TC_B_BLOCK_BYTES, 2 * TC_B_BLOCK_BYTES,
                              smem_u32(&S.b_full[stage]), keep);
                if (++stage == BS) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 3 && lane == 0) {
        // ===== feature producer: one hand tile, read once =====
        const uint64_t once = l2_policy_evict_first();
        uint32_t phase = 0;
        bool ok = true;
        for (int m_tile = blockIdx.x; m_tile < m_tiles && ok; m_tile += gridDim.x) {
            const unsigned char* fsrc = featp + (size_t)m_tile * TC_A_TILE_BYTES;
            for (int c = 0; c < TC_K_CHUNKS; ++c) {
                if (!(ok = mbar_wait(smem_u32(&S.a_empty[c]), phase ^ 1, abort_flag))) break;
                mbar_expect_tx(smem_u32(&S.a_full[c]), TC_A_STAGE_BYTES);
                bulk_g2s_hint(smem_u32(S.a[c]), fsrc + (size_t)c * TC_A_STAGE_BYTES, TC_A_STAGE_BYTES, smem_u32(&S.a_full[c]), once);
            }
            phase ^= 1;
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        uint32_t stage = 0, phase = 0, a_phase = 0, acc = 0, acc_phase = 0;
        bool ok = true;
        const uint64_t a_base = umma_desc(smem_u32(S.a[0]), TC_LBO, TC_SBO);
        const uint64_t b_base = umma_desc(smem_u32(S.b[0]), TC_LBO, TC_SBO);
        for (int m_tile = blockIdx.x; m_tile < m_tiles && ok; m_tile += gridDim.x) {
            for (int n_tile = 0; n_tile < TC_N_TILES && ok; ++n_tile) {
                if (!(ok = mbar_wait(smem_u32(&S.acc_empty[acc]), acc_phase ^ 1, abort_flag))) break;
                tc_fence_after();
                const uint32_t d_tmem = tmem + acc * TC_N;
#pragma unroll
                for (int c = 0; c < TC_K_CHUNKS; ++c) {
                    if (n_tile == 0 && !(ok = mbar_wait(smem_u32(&S.a_full[c]), a_phase, abort_flag))) break;
                    if (!(ok = mbar_wait(smem_u32(&S.b_full[stage]), phase, abort_flag))) break;
                    tc_fence_after();
                    const uint64_t b_st = b_base + (uint64_t)((stage * 2 * TC_B_BLOCK_BYTES) >> 4);
#pragma unroll
                    for (int j = 0; j < TC_K_CHUNK / 16; ++j) {
                        const uint64_t a_hi = a_base + (uint64_t)((c * TC_A_STAGE_BYTES + j * 2 * (int)TC_LBO) >> 4);
                        const uint64_t a_lo = a_hi + (uint64_t)(TC_A_BLOCK_BYTES >> 4);
                        const uint64_t b_hi = b_st + (uint64_t)((j * 2 * (int)TC_LBO) >> 4);
                        const uint64_t b_lo = b_hi + (uint64_t)(TC_B_BLOCK_BYTES >> 4);
                        umma_f16(d_tmem, a_hi, b_hi, IDESC, (c | j) ? 1u : 0u);
                        if (products == 3) {
                            umma_f16(d_tmem, a_lo, b_hi, IDESC, 1);
                            umma_f16(d_tmem, a_hi, b_lo, IDESC, 1);
                        }
                    }
                    tc_commit(smem_u32(&S.b_empty[stage]));            // frees the basis slot when these MMAs retire
                    if (n_tile == TC_N_TILES - 1) tc_commit(smem_u32(&S.a_empty[c]));   // ... and the feature chunk after its last use
                    if (++stage == BS) { stage = 0; phase ^= 1; }
                }
                if (!ok) break;
                tc_commit(smem_u32(&S.acc_full[acc]));
                if (++acc == MRES_ACC) { acc = 0; acc_phase ^= 1; }
            }
            a_phase ^= 1;
        }
    } else if (warp >= 4) {
        // ===== epilogue (as above; v_template from shared memory, one broadcast read per column) =====
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int j_begin = half == 0 ? 0 : 3, j_end = half == 0 ? 3 : TC_N / 32;
        uint32_t acc = 0, acc_phase = 0;
        bool ok = true;
        for (int m_tile = blockIdx.x; m_tile < m_tiles && ok; m_tile += gridDim.x) {
            const long long group = (long long)m_tile * (TC_M / 32) + q;
            const bool live = group * 32 < B;
            for (int n_tile = 0; n_tile < TC_N_TILES; ++n_tile) {
                ok = __all_sync(0xffffffffu, mbar_wait(smem_u32(&S.acc_full[acc]), acc_phase, abort_flag));
                if (!ok) break;
                tc_fence_after();
#pragma unroll 1
                for (int j = j_begin; j < j_end; ++j) {
                    float v[32];
                    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + acc * TC_N + j * 32, v);
                    if (j == j_end - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&S.acc_empty[acc]));
                    }
                    const int col0 = n_tile * TC_N + j * 32;
                    if (live) {
                        float* dst = v_posed_t + ((size_t)group * SK_NCOORD + col0) * 32 + lane;
                        const float4* tp = reinterpret_cast<const float4*>(S.tmpl + col0);
#pragma unroll
                        for (int c4 = 0; c4 < 8; ++c4) {
                            const float4 t4 = tp[c4];
                            const int c = c4 * 4;
                            if (col0 + c < SK_NCOORD) {               // SK_NCOORD is a multiple of 4
                                __stcs(dst + (c + 0) * 32, fmaf(v[c + 0], out_scale, t4.x));
                                __stcs(dst + (c + 1) * 32, fmaf(v[c + 1], out_scale, t4.y));
                                __stcs(dst + (c + 2) * 32, fmaf(v[c + 2], out_scale, t4.z));
                                __stcs(dst + (c + 3) * 32, fmaf(v[c + 3], out_scale, t4.w));
                            }
                        }
                    }
                }
                if (++acc == MRES_ACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    }
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, TMEM_COLS);
    if (threadIdx.x == 0 && S.abort_flag) __trap();
}


// ================================================================= backward contraction
//   dfeat[h][n] = sum_k dv_posed[h][k] * basis[n][k],  n < 145 (+ pad to 160), K = 2336
// M = 2 x 128 hands per CTA pass (two accumulators of 160 TMEM columns), both operands stream
// through a 3-stage TMA/mbarrier ring (per K chunk of 32: 2 x 16 KB of dv tiles + 20 KB of basis),
// bf16 hi + mid split on both sides, three products (hi*hi + mid*hi + hi*mid) in fp32 TMEM
// accumulators: bf16 keeps fp32's exponent range, so upstream gradients of any magnitude need no
// scaling, and 16 significand bits per operand bound the error at ~2e-5 relative.
constexpr int BW_STAGES = 3;
constexpr int BW_MT = 2;                                           // hand tiles per CTA pass
constexpr int BW_STAGE_BYTES = BW_MT * TCB_A_CHUNK_BYTES + TCB_B_CHUNK_BYTES;   // 52 KB
constexpr uint32_t IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

struct TcBwdShared {
    alignas(128) unsigned char st[BW_STAGES][BW_STAGE_BYTES];       // [A tile 0 (hi,mid)][A tile 1 (hi,mid)][B (hi,mid)]
    alignas(16) float stage[EPI_WARPS][32][33];
    alignas(8) unsigned long long full[BW_STAGES], empty[BW_STAGES];
    unsigned long long acc_full, acc_empty;
    uint32_t tmem_base;
    int abort_flag;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
blend_tc_backward_kernel(const unsigned char* __restrict__ basis_bw, const unsigned char* __restrict__ dvp,
                         float* __restrict__ dfeat, int B, int m_tiles, int hand_minor, int ksplit, size_t part_stride) {
    // declared with its alignment and used WITHOUT integer arithmetic on the address: rounding the pointer up through uintptr_t
    // made the compiler lose the shared address space — every access became a generic LD / ST [profiles/r2]
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcBwdShared& S = *reinterpret_cast<TcBwdShared*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work item w = (pass p = w / ksplit, K range w % ksplit): small batches cut the K loop so that every SM has a
    // range; range r accumulates chunks [74 r / ksplit, 74 (r + 1) / ksplit) into its own dfeat copy
    const int passes = (m_tiles + BW_MT - 1) / BW_MT;
    const int nwork = passes * ksplit;

    if (threadIdx.x == 0) {
        for (int s = 0; s < BW_STAGES; ++s) { mbar_init(smem_u32(&S.full[s]), 1); mbar_init(smem_u32(&S.empty[s]), 1); }
        mbar_init(smem_u32(&S.acc_full), 1);
        mbar_init(smem_u32(&S.acc_empty), EPI_WARPS);
        S.abort_flag = 0;
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(&S.tmem_base), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    volatile int* abort_flag = &S.abort_flag;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        uint32_t stage = 0, phase = 0;
        bool ok = true;
        for (int w = blockIdx.x; w < nwork && ok; w += gridDim.x) {
            const int p = w / ksplit, r = w - p * ksplit;
            const int nmt = (m_tiles - p * BW_MT) < BW_MT ? (m_tiles - p * BW_MT) : BW_MT;
            for (int kc = TCB_K_CHUNKS * r / ksplit; kc < TCB_K_CHUNKS * (r + 1) / ksplit; ++kc) {
                if (!(ok = mbar_wait(smem_u32(&S.empty[stage]), phase ^ 1, abort_flag))) break;
                mbar_expect_tx(smem_u32(&S.full[stage]), nmt * TCB_A_CHUNK_BYTES + TCB_B_CHUNK_BYTES);
                const uint32_t dst = smem_u32(S.st[stage]);
                for (int mt = 0; mt < nmt; ++mt)
                    bulk_g2s(dst + mt * TCB_A_CHUNK_BYTES,
                             dvp + (size_t)(p * BW_MT + mt) * TCB_A_TILE_BYTES + (size_t)kc * TCB_A_CHUNK_BYTES,
                             TCB_A_CHUNK_BYTES, smem_u32(&S.full[stage]));
                bulk_g2s(dst + BW_MT * TCB_A_CHUNK_BYTES, basis_bw + (size_t)kc * TCB_B_CHUNK_BYTES, TCB_B_CHUNK_BYTES,
                         smem_u32(&S.full[stage]));
                if (++stage == BW_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        uint32_t stage = 0, phase = 0, pass_phase = 0;
        bool ok = true;
        const uint64_t base = umma_desc(smem_u32(S.st[0]), TC_LBO, TC_SBO);
        for (int w = blockIdx.x; w < nwork && ok; w += gridDim.x) {
            const int p = w / ksplit, r = w - p * ksplit;
            const int nmt = (m_tiles - p * BW_MT) < BW_MT ? (m_tiles - p * BW_MT) : BW_MT;
            const int kc0 = TCB_K_CHUNKS * r / ksplit, kc1 = TCB_K_CHUNKS * (r + 1) / ksplit;
            if (!(ok = mbar_wait(smem_u32(&S.acc_empty), pass_phase ^ 1, abort_flag))) break;   // epilogue drained the accumulators
            tc_fence_after();
            for (int kc = kc0; kc < kc1 && ok; ++kc) {
                if (!(ok = mbar_wait(smem_u32(&S.full[stage]), phase, abort_flag))) break;
                tc_fence_after();
                const uint64_t sbase = base + (uint64_t)((stage * BW_STAGE_BYTES) >> 4);
                const uint64_t b_hi0 = sbase + (uint64_t)((BW_MT * TCB_A_CHUNK_BYTES) >> 4);
#pragma unroll
                for (int mt = 0; mt < BW_MT; ++mt) {
                    if (mt < nmt) {
#pragma unroll
                        for (int j = 0; j < TC_K_CHUNK / 16; ++j) {
                            const uint64_t a_hi = sbase + (uint64_t)((mt * TCB_A_CHUNK_BYTES + j * 2 * (int)TC_LBO) >> 4);
                            const uint64_t a_mid = a_hi + (uint64_t)(TC_A_BLOCK_BYTES >> 4);
                            const uint64_t b_hi = b_hi0 + (uint64_t)((j * 2 * (int)TC_LBO) >> 4);
                            const uint64_t b_mid = b_hi + (uint64_t)(TC_B_BLOCK_BYTES >> 4);
                            const uint32_t d = tmem + mt * TC_N;
                            umma_f16(d, a_hi, b_hi, IDESC_BF16, (kc != kc0 || j) ? 1u : 0u);
                            umma_f16(d, a_mid, b_hi, IDESC_BF16, 1);
                            umma_f16(d, a_hi, b_mid, IDESC_BF16, 1);
                        }
                    }
                }
                tc_commit(smem_u32(&S.empty[stage]));
                if (++stage == BW_STAGES) { stage = 0; phase ^= 1; }
            }
            if (!ok) break;
            tc_commit(smem_u32(&S.acc_full));
            pass_phase ^= 1;
        }
    } else if (warp >= 4) {
        // ===== epilogue: warps 4-7 drain hand tile 0, warps 8-11 hand tile 1 =====
        const int q = warp & 3;
        const int mt = (warp - 4) >> 2;
        float (*buf)[33] = S.stage[warp - 4];
        uint32_t pass_phase = 0;
        bool ok = true;
        for (int w = blockIdx.x; w < nwork && ok; w += gridDim.x) {
            const int p = w / ksplit;
            float* dfeat_r = dfeat + (size_t)(w - p * ksplit) * part_stride;
            ok = __all_sync(0xffffffffu, mbar_wait(smem_u32(&S.acc_full), pass_phase, abort_flag));
            if (!ok) break;
            tc_fence_after();
            const int m_tile = p * BW_MT + mt;
            const int row0 = m_tile * TC_M + q * 32;
            const bool active = m_tile < m_tiles;
#pragma unroll 1
            for (int j = 0; j < TC_N / 32; ++j) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + mt * TC_N + j * 32, v);
                if (j == TC_N / 32 - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&S.acc_empty));
                }
                if (hand_minor) {                               // dfeat_t[group][column][32]: a lane quarter is a hand group
                    const long long group = (long long)m_tile * (TC_M / 32) + q;
                    if (active && group * 32 < B) {
                        float* dst = dfeat_r + ((size_t)group * TC_N + j * 32) * 32 + lane;
#pragma unroll
                        for (int c = 0; c < 32; ++c) dst[c * 32] = v[c];
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) buf[lane][c] = v[c];
                    __syncwarp();
                    const int col = j * 32 + lane;
                    if (active && col < FEAT_K) {
                        float* dst = dfeat_r + (size_t)row0 * FEAT_K + col;
                        const int nrow = B - row0 < 32 ? B - row0 : 32;
#pragma unroll 8
                        for (int rr = 0; rr < 32; ++rr)
                            if (rr < nrow) dst[(size_t)rr * FEAT_K] = buf[rr][lane];
                    }
                    __syncwarp();
                }
            }
            pass_phase ^= 1;
        }
    }
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, TMEM_COLS);
    if (threadIdx.x == 0 && S.abort_flag) __trap();
}

}  // namespace

// ---------------------------------------------------------------- host side
static size_t tc_fwd_bytes() { return align256(sizeof(TcBlobHeader)) + align256((size_t)TC_N_TILES * TC_B_TILE_BYTES); }
size_t blend_tc_blob_bytes() { return tc_fwd_bytes() + align256(TCB_B_BYTES); }

// basis [FEAT_K][2334] fp32 -> per n-tile, per K chunk, {hi, lo} blocks in the UMMA canonical
// K-major layout [row-group][k-group][8 rows][8 halves]; pre-scaled by a power of two.  Columns are in
// the skinning block order: coord_map[column] = original coordinate (vertex*3+c) or -1 (padding).
void blend_tc_pack(const float* basis, const int32_t* coord_map, void* host_blob_tc) {
    unsigned char* out = reinterpret_cast<unsigned char*>(host_blob_tc);
    memset(out, 0, blend_tc_blob_bytes());
    TcBlobHeader* H = reinterpret_cast<TcBlobHeader*>(out);
    float mx = 0.f;
    for (int k = 0; k < TC_K_REAL; ++k)
        for (int n = 0; n < NVC; ++n) mx = fmaxf(mx, fabsf(basis[(size_t)k * NVC + n]));
    int e = 13;                                               // scale 2^e: largest with max * 2^e <= 2^11
    if (mx > 0.f) { e = 11 - (int)ceilf(log2f(mx)); if (e > 24) e = 24; if (e < -8) e = -8; }
    const float sb = ldexpf(1.f, e);
    H->basis_scale_log2 = e;
    H->feat_scale_log2 = TC_FEAT_SCALE_LOG2;
    __half* dst = reinterpret_cast<__half*>(out + align256(sizeof(TcBlobHeader)));
    for (int nt = 0; nt < TC_N_TILES; ++nt)
        for (int c = 0; c < TC_K_CHUNKS; ++c)
            for (int r = 0; r < TC_N; ++r)
                for (int kk = 0; kk < TC_K_CHUNK; ++kk) {
                    const int n = coord_map[nt * TC_N + r], k = c * TC_K_CHUNK + kk;     // column -> original coordinate
                    float x = (n >= 0 && k < TC_K_REAL) ? basis[(size_t)k * NVC + n] * sb : 0.f;
                    const __half hi = __float2half_rn(x);
                    const __half lo = __float2half_rn(x - __half2float(hi));
                    const size_t blk = ((size_t)nt * TC_K_CHUNKS + c) * 2;
                    const size_t in = (((size_t)(r >> 3) * (TC_K_CHUNK / 8) + (kk >> 3)) * 8 + (r & 7)) * 8 + (kk & 7);
                    dst[(blk + 0) * (TC_B_BLOCK_BYTES / 2) + in] = hi;
                    dst[(blk + 1) * (TC_B_BLOCK_BYTES / 2) + in] = lo;
                }
    // backward B operand: basis[n][k] as bf16 hi + mid (unscaled), rows n = feature (160, zero beyond 145),
    // per K chunk of 32 block-order coordinates: [74][split 2][row-group 20][k-group 4][8 rows][8 bf16]
    __nv_bfloat16* bw = reinterpret_cast<__nv_bfloat16*>(out + tc_fwd_bytes());
    for (int kc = 0; kc < TCB_K_CHUNKS; ++kc)
        for (int r = 0; r < TC_N; ++r)
            for (int kk = 0; kk < TC_K_CHUNK; ++kk) {
                const int kcol = kc * TC_K_CHUNK + kk;
                const int k = kcol < SK_TMPL_PAD ? coord_map[kcol] : -1;
                const float x = (r < TC_K_REAL && k >= 0) ? basis[(size_t)r * NVC + k] : 0.f;
                const __nv_bfloat16 hi = __float2bfloat16_rn(x);
                const __nv_bfloat16 mid = __float2bfloat16_rn(x - __bfloat162float(hi));
                const size_t in = (((size_t)(r >> 3) * (TC_K_CHUNK / 8) + (kk >> 3)) * 8 + (r & 7)) * 8 + (kk & 7);
                bw[((size_t)kc * 2 + 0) * (TC_B_BLOCK_BYTES / 2) + in] = hi;
                bw[((size_t)kc * 2 + 1) * (TC_B_BLOCK_BYTES / 2) + in] = mid;
            }
}

int launch_blend_tc_backward(const void* blob, const unsigned char* dvp, float* dfeat, int B, int hand_minor, int ksplit,
                             size_t part_stride, cudaStream_t s) {
    if (B <= 0) return 0;
    static SmemAttrOnce once;
    const size_t smem = sizeof(TcBwdShared) + 128;
    if (int arc = ensure_dyn_smem(once, blend_tc_backward_kernel, smem)) return arc;
    const BlobLayout L = blob_layout();
    const unsigned char* tc = blob_ptr<unsigned char>(blob, L.total);
    const int m_tiles = (B + TC_M - 1) / TC_M;
    const int passes = (m_tiles + BW_MT - 1) / BW_MT;
    if (ksplit < 1 || ksplit > TCB_K_CHUNKS) return MB_E_RANGE;
    const long long nwork = (long long)passes * ksplit;
    blend_tc_backward_kernel<<<nwork < NUM_SMS ? (int)nwork : NUM_SMS, TC_THREADS, smem, s>>>(tc + tc_fwd_bytes(), dvp, dfeat, B, m_tiles,
                                                                                          hand_minor, ksplit, part_stride);
    return cuda_rc();
}

int launch_blend_tc_forward(const void* blob, const unsigned char* featp, float* v_posed_t, int B, int mode, cudaStream_t s) {
    if (B <= 0) return 0;
    static SmemAttrOnce once;
    const size_t smem = sizeof(TcShared) + 128;
    if (int arc = ensure_dyn_smem(once, blend_tc_forward_kernel, smem)) return arc;
    const BlobLayout L = blob_layout();
    const unsigned char* tc = blob_ptr<unsigned char>(blob, L.total);
    const int m_tiles = (B + TC_M - 1) / TC_M;
    const long long total = (long long)TC_N_TILES * m_tiles;
    const int grid = (int)(total < NUM_SMS ? total : NUM_SMS);      // >= 15: every n-tile has at least one CTA
    const float* tmpl = blob_ptr<float>(blob, L.sk_tmpl);
    if (m_tiles >= TC_MRES_MIN_TILES) {             // large batch: hand tile resident, basis streamed (one CTA per hand tile)
        static SmemAttrOnce once2;
        const size_t smem2 = sizeof(TcSharedM) + 128;
        if (int arc = ensure_dyn_smem(once2, blend_tc_forward_mres_kernel, smem2)) return arc;
        blend_tc_forward_mres_kernel<<<m_tiles < NUM_SMS ? m_tiles : NUM_SMS, TC_THREADS, smem2, s>>>(
            reinterpret_cast<const TcBlobHeader*>(tc), tc + align256(sizeof(TcBlobHeader)), tmpl, featp, v_posed_t, B, m_tiles,
            mode == MB_MODE_F16X3 ? 3 : 1);
        return cuda_rc();
    }
    blend_tc_forward_kernel<<<grid, TC_THREADS, smem, s>>>(reinterpret_cast<const TcBlobHeader*>(tc),
                                                           tc + align256(sizeof(TcBlobHeader)), tmpl, featp, v_posed_t, B, m_tiles,
                                                           mode == MB_MODE_F16X3 ? 3 : 1);
    return cuda_rc();
}

}  // namespace mb
